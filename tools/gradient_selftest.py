import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import niftymatch_b200 as nm
lib = nm.load()
for seed in range(4):
    m = C.c_longlong(-1)
    rc = lib.nm_selftest_gradient(1 << 28, seed, C.byref(m))
    print("seed", seed, "rc", rc, "mismatches", m.value)
