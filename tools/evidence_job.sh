bash tools/gpu_job.sh tests bench configs launches
KEEP_REP=1 bash tools/gpu_job.sh "ncu:fused_d12:jne_run_kernel:--dim,12,--T,10000,--n,133200"
python tools/ncu_traffic.py gpurun_out/prof_fused_d12.ncu-rep --dim 12 --T 10000 --n 133200 --out gpurun_out/traffic_fused_d12_jne2.json; rm -f gpurun_out/prof_fused_d12.ncu-rep
bash tools/gpu_job.sh "ncu:m0_d12:jne_run_kernel:--dim,12,--T,10000,--n,133200,--models,0" "ncu:m4_d12:jne_run_kernel:--dim,12,--T,10000,--n,133200,--models,4" "ncu:fused_d8:jne_run_kernel:--dim,8,--T,10000,--n,133200" "ncu:lane_d5:jne_lane_moments_kernel:--dim,5,--T,5000,--n,227328" "ncu:lane_d1:jne_lane_moments_kernel:--dim,1,--T,10000,--n,151552" "ncu:group_d9:jne_group_moments_kernel:--dim,9,--T,10000,--n,118400"
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_target.py 2>&1 | tail -n 5 | tee gpurun_out/sanitizer_memcheck.txt
timeout 500 compute-sanitizer --tool racecheck python tools/sanitize_target.py 2>&1 | tail -n 5 | tee gpurun_out/sanitizer_racecheck.txt
bash tools/gpu_job.sh "py:rng_moments:--n-log2,29"
