"""Diagnostic: the in-process multi-device context while ANOTHER process keeps one of its GPUs busy (as under torchrun)."""
import subprocess, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import johansen_null_eigenspectra_b200 as jne
nd = torch.cuda.device_count()
R = 133200
seeds = np.arange(1, R * nd + 1, dtype=np.uint32)
def run(tag):
    eng = jne.Engine(None)
    buf = np.empty((seeds.size, 62))
    try:
        eng.eigs_batch_multi(range(5), 12, 10000, seeds[: 4096 * nd])
        t0 = time.time()
        eng.eigs_batch_multi(range(5), 12, 10000, seeds, out=buf)
        print(tag, "ok", f"{5 * seeds.size / (time.time() - t0) / 1e6:.1f} M runs/s", "finite:", bool(np.isfinite(buf).all()), flush=True)
    except Exception as e:
        print(tag, "FAILED", e, "non-finite rows:", int((~np.isfinite(buf)).any(axis=1).sum()),
              "first bad rows:", np.nonzero((~np.isfinite(buf)).any(axis=1))[0][:10], flush=True)
    eng.close()
one = jne.Engine([0]); one.eigs_batch_multi(range(5), 12, 10000, seeds[:8192]); 
run("alone")
spin = subprocess.Popen([sys.executable, "-c", "import torch,time\nx=torch.randn(8192,8192,device='cuda:%d')\nt=time.time()\nwhile time.time()-t<25: y=x@x; torch.cuda.synchronize()" % (nd - 1)])
time.sleep(6)
run("with another process computing on the last GPU")
run("again")
spin.wait()
run("after")
