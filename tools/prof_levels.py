"""Profiling driver (not a test): every blur level of the octave-0 schedule at MNI size through the stage-level
C-ABI call, cold (256 MiB L2 flush before each level) and in a warm chain (a -> b -> a ...), for several
tuning configurations given as arguments ("K=V,K=V" each; "" = defaults).  One Engine per configuration."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("3d_sift_cuda_b200")
if os.environ.get("PROF_LIB"):      # A/B runs of two builds on the same box
    pkg.api.library_path = lambda: os.path.join(os.path.dirname(pkg.api.__file__), os.environ["PROF_LIB"])
X, Y, Z = [int(v) for v in os.environ.get("PROF_SHAPE", "182,218,182").split(",")]
pitch = (X + 7) // 8 * 8
rng = np.random.default_rng(0)
a = torch.zeros((Z, Y, pitch), dtype=torch.float32, device="cuda")
a[:, :, :X] = torch.from_numpy(rng.random((Z, Y, X), dtype=np.float32)).cuda()
b, tmp, dog = torch.zeros_like(a), torch.zeros_like(a), torch.zeros_like(a)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
sigmas = [1.5199, 1.2263, 1.5450, 1.9466, 2.4525, 3.0900]
if os.environ.get("PROF_SIGMAS"):
    sigmas = [float(v) for v in os.environ["PROF_SIGMAS"].split(",")]
REPS = int(os.environ.get("PROF_REPS", "10"))
N0 = X * Y * Z
PEAK = 6554.6
configs = sys.argv[1:] or [""]
for cfg in configs:
    saved = {}
    for kv in [c for c in cfg.split(",") if c]:
        k, v = kv.split("=")
        saved[k] = os.environ.get(k)
        os.environ[k] = v
    e = pkg.Engine(0)
    st = torch.cuda.ExternalStream(e.stream)
    cold, warm = [], []
    for s in sigmas:
        taps = pkg.gaussian_taps(s)
        for _ in range(0 if os.environ.get("NCU_ONE") else 3):
            e.blur3d(a, tmp, b, X, taps, dog)
        e.sync()
        ms = 0.0
        with torch.cuda.stream(st):
            for _ in range(REPS):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st); e.blur3d(a, tmp, b, X, taps, dog); e1.record(st)
                e.sync()
                ms += e0.elapsed_time(e1)
        cold.append(ms / REPS * 1e3)
        N = 0 if os.environ.get("NCU_ONE") else 20
        if N == 0:
            warm.append(float('nan')); continue
        with torch.cuda.stream(st):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for i in range(N):
                if i % 2 == 0: e.blur3d(a, tmp, b, X, taps, dog)
                else: e.blur3d(b, tmp, a, X, taps, dog)
            e1.record(st)
        e.sync()
        warm.append(e0.elapsed_time(e1) * 1e3 / N)
    e.close()
    for k, v in saved.items():
        if v is None: del os.environ[k]
        else: os.environ[k] = v
    ntaps = [len(pkg.gaussian_taps(s)) for s in sigmas]
    print("[%s]" % (cfg or "default"))
    print("  cold us : " + " ".join("%2dt %6.1f" % (n, t) for n, t in zip(ntaps, cold)) + " | sum %.1f" % sum(cold))
    print("  cold frac: " + " ".join("%2dt %6.3f" % (n, 12.0 * N0 / (t * 1e-6) / 1e9 / PEAK) for n, t in zip(ntaps, cold)))
    print("  warm us : " + " ".join("%2dt %6.1f" % (n, t) for n, t in zip(ntaps, warm)) + " | sum %.1f" % sum(warm), flush=True)
