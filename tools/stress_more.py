"""More seeds of the randomized geometry / sensor-constant stress test than the suite runs (tests/test_gpu_parity.py
::test_random_geometry_and_sensor_constants): every engine, warps per CTA and lane layout against the oracle."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as T
from oracle import orc
gpu = importlib.import_module("micro-quad-slam_b200"); gpu.init(0)
o = orc.Oracle()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
t0 = time.time(); bad = 0
for seed in range(lo, hi):
    try:
        T.test_random_geometry_and_sensor_constants(gpu, o, seed)
    except AssertionError as e:
        bad += 1
        print(f"seed {seed}: FAIL {str(e)[:300]}", flush=True)
print(f"seeds {lo}..{hi - 1}: {hi - lo - bad} ok, {bad} failed in {time.time() - t0:.0f} s")
