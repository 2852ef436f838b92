"""Profiling driver (not a test): the CLI's list mode on N gzip-compressed int16 MNI-sized volumes
(featExtract -l list): wall time per volume with file decoding, extraction and feature-file writing, for
S3D_CLI_THREADS-style comparisons use `taskset -c 0` to see the single-host-thread figure."""
import gzip, importlib, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("3d_sift_cuda_b200")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "3d_sift_cuda_b200", "featExtract")
with tempfile.TemporaryDirectory() as d:
    lines = []
    for k in range(min(N, 8)):
        v = np.rint(pkg.phantom.brain_phantom((182, 218, 182), 1 + k, 400) * 128).astype(np.int16)
        pkg.phantom.write_nifti(os.path.join(d, "p%d.nii" % k), v, dtype=np.int16)
        with open(os.path.join(d, "p%d.nii" % k), "rb") as f, gzip.open(os.path.join(d, "p%d.nii.gz" % k), "wb", compresslevel=6) as g:
            g.write(f.read())
        os.remove(os.path.join(d, "p%d.nii" % k))
    for k in range(N):
        lines.append("%s %s" % (os.path.join(d, "p%d.nii.gz" % (k % 8)), os.path.join(d, "out%d.key" % k)))
    open(os.path.join(d, "list.txt"), "w").write("\n".join(lines) + "\n")
    for prefix, label in ((["taskset", "-c", "0"], "one host core"), ([], "all host cores")):
        t0 = time.perf_counter()
        r = subprocess.run(prefix + [exe, "-l", os.path.join(d, "list.txt")], capture_output=True, text=True, env=dict(os.environ, S3D_CLI_TIMING="1"))
        dt = time.perf_counter() - t0
        ok = r.returncode == 0 and all(os.path.getsize(os.path.join(d, "out%d.key" % k)) > 1000 for k in range(N))
        print("featExtract -l, %d x 182x218x182 int16 .nii.gz, %s: %.2f s = %.1f ms per volume (%s)" % (N, label, dt, 1e3 * dt / N, "ok" if ok else "FAILED: " + r.stdout[-300:]))
        print("   " + r.stderr.strip()[-400:])
