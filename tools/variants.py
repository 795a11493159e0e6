"""Build-time ablations of the hot kernels (the lever table of DESIGN.md section 5).
  python tools/variants.py build     HERE (nvcc cross-compiles): libjne_var_<name>.so per variant, in parallel
  python tools/variants.py run       on the GPU box: times every variant on the metric's configuration and on c2
Each variant is the production library with one -D switch; none of them is loaded by the product."""
import json, os, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "johansen_null_eigenspectra_b200"
VARIANTS = {
    "base": [],
    "nobm": ["-DJNE_EXP_NOBM"],                    # generator only, no normal transform (NOT a valid stream)
    "norng": ["-DJNE_EXP_NORNG"],                  # no generator at all (NOT a valid stream)
    "seglen8": ["-DJNE_EXP_SEGLEN8"],              # tensor family: segments of whole 8-step blocks instead of whole epochs (NOT a valid stream): cost of the longer masked tail
    "noreseed": ["-DJNE_EXP_NORESEED"],            # tensor family: substreams keyed once per run (NOT a valid stream): cost of the in-loop key generation
    "seglen8_noreseed": ["-DJNE_EXP_SEGLEN8", "-DJNE_EXP_NORESEED"],
    "minb4": ["-DJNE_MULTI_MINB=4"],               # fused tensor kernel at 4 CTAs per SM (116 registers: no spill of the substream states)
    "lane6_reg_states": ["-DJNE_LANE_REG_STATES=6"],     # lane family dim 6: generator states in registers (spills 72 bytes)
    "lane_minb_hi": ["-DJNE_LANE_MINB3=5", "-DJNE_LANE_MINB4=4"],   # lane family dims 3 / 4 at 102 / 128 registers (the substream states spill)
    "group78": ["-DJNE_EXP_GROUP_78"],             # dims 7, 8 on the group kernel (2 lanes x 4 rows), dim 10 as 5 x 2
    "lane_nopipe": ["-DJNE_LANE_PIPELINE=0"],      # lane family without the software pipeline (generate a block, then consume it)
    "lane_nopipe_minb3": ["-DJNE_LANE_PIPELINE=0", "-DJNE_LANE_MINB5=3"],   # ... and dim 5 at 168 registers / 3 CTAs per SM
}
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-pthread"]


def lib(name):
    return PKG / f"libjne_var_{name}.so"


def build(names):
    procs = {}
    for name in names:
        cmd = ["nvcc", *FLAGS, *VARIANTS[name], "-I", str(ROOT / "include"), "-o", str(lib(name)),
               *[str(PKG / "csrc" / f) for f in ("jne_api.cu", "jne_dat.cpp", "jne_host.cpp")]]
        procs[name] = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    for name, p in procs.items():
        out, _ = p.communicate()
        print(name, "ok" if p.returncode == 0 else "FAILED\n" + out[-2000:])


WORKER = r'''
import json, sys, torch
sys.path.insert(0, ".")
import johansen_null_eigenspectra_b200 as jne
eng = jne.Engine([0]); st = torch.cuda.current_stream()
def t(models, dim, T, n, reps=3):
    seeds = torch.arange(1, n + 1, dtype=torch.int32, device="cuda")
    out = torch.empty((n, sum(jne.num_eigs(m, dim) for m in models)), dtype=torch.float64, device="cuda")
    best = 1e30
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        if len(models) == 1: eng.eigs_batch_device(models[0], dim, T, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream)
        else: eng.eigs_batch_multi_device(models, dim, T, seeds.data_ptr(), n, out.data_ptr(), st.cuda_stream)
        e1.record(st); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return n / best / 1e3      # M seeds/s
res = {"fused_d12": t(range(5), 12, 10000, 133200), "m0_d12": t([0], 12, 10000, 133200), "m4_d12": t([4], 12, 10000, 133200),
       "fused_d9": t(range(5), 9, 10000, 118400), "fused_d10": t(range(5), 10, 10000, 118400), "fused_d11": t(range(5), 11, 10000, 133200), "fused_d8": t(range(5), 8, 10000, 133200), "fused_d7": t(range(5), 7, 10000, 133200),
       "fused_d5_T5000": t(range(5), 5, 5000, 1 << 20), "m0_d5_T5000": t([0], 5, 5000, 1 << 20), "m4_d5_T5000": t([4], 5, 5000, 1 << 20),
       "fused_d1": t(range(5), 1, 10000, 1 << 20), "fused_d2": t(range(5), 2, 10000, 1 << 20), "fused_d3": t(range(5), 3, 10000, 1 << 20),
       "fused_d4": t(range(5), 4, 10000, 1 << 20), "fused_d6": t(range(5), 6, 10000, 1 << 19)}
print(json.dumps(res))
'''


def run(names):
    print("# M seeds/s per variant (device-resident, best of 3)")
    for name in names:
        if not lib(name).exists():
            print(name, "not built"); continue
        r = subprocess.run([sys.executable, "-c", WORKER], cwd=ROOT, env=dict(os.environ, JNE_LIBRARY=str(lib(name))),
                           capture_output=True, text=True)
        if r.returncode != 0:
            print(name, "FAILED", r.stderr[-1500:]); continue
        res = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{name:14s} " + "  ".join(f"{k} {v:8.3f}" for k, v in res.items()), flush=True)


if __name__ == "__main__":
    names = [a for a in sys.argv[2:] if a in VARIANTS] or list(VARIANTS)
    {"build": build, "run": run}[sys.argv[1]](names)
