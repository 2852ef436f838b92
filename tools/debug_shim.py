"""Debug driver (GPU box): which stage of the shim build (oracle/_ref/featExtract_ref_s3d) departs from the CPU path."""
import importlib, os, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3d_sift_cuda_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "featExtract_ref")
S3D = os.path.join(ROOT, "oracle", "_ref", "featExtract_ref_s3d")
d = tempfile.mkdtemp()
vol = pkg.phantom.blob_phantom((56, 60, 52), 31, 45)
nii = os.path.join(d, "in.nii")
pkg.phantom.write_nifti(nii, vol)
subprocess.run([REF, nii, "cpu.key"], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
a = open(os.path.join(d, "cpu.key"), "rb").read()
for cpu in ["blur,dog,sub,det", "dog,sub,det", "blur,sub,det", "blur,dog,det", "blur,dog,sub", ""]:
    env = dict(os.environ, S3D_SHIM_CPU=cpu)
    r = subprocess.run([S3D, "-d0", nii, "s3d.key"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)
    b = open(os.path.join(d, "s3d.key"), "rb").read() if r.returncode == 0 else b""
    print("cpu stages [%s]: rc %d, %d vs %d lines, equal %s" % (cpu, r.returncode, a.count(b"\n"), b.count(b"\n"), a == b), flush=True)
    if r.returncode != 0: print(r.stdout.decode(errors="replace")[-500:])
