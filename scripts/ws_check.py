"""Quick GPU check of the warp-specialised pass kernel: bit-exactness against numpy / the oracle + timing."""
import ctypes, os, sys, time
import numpy as np
import torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import compute_b200 as cb
import gpu_api, oracle

def stats():
    r, f = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    cb.lib().bcb_sort_speculation_stats(cb.command_queue().handle, ctypes.byref(r), ctypes.byref(f))
    return r.value, f.value

rng = np.random.default_rng(1)
ok = True
for n in ((1 << 23), (1 << 23) + 1, (1 << 24) + 777, 3 * (1 << 24) + 12345):
    k = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    out = gpu_api.radix_sort(k)
    good = np.array_equal(out, np.sort(k))
    print("u32 asc", n, "OK" if good else "MISMATCH", stats(), flush=True)
    ok &= good
for dtype, desc in (("int", True), ("float", False), ("ulong", False), ("double", False), ("long", True), ("uint", True)):
    n = (1 << 23) + 4321
    npdt = {"int": np.int32, "float": np.float32, "ulong": np.uint64, "double": np.float64, "long": np.int64, "uint": np.uint32}[dtype]
    k = rng.integers(0, 2**(8 * np.dtype(npdt).itemsize), size=n, dtype=np.uint64).astype(np.dtype(npdt).str.replace("f", "u").replace("i", "u")).view(npdt)
    out = gpu_api.radix_sort(k, desc)
    good = out.tobytes() == oracle.radix_sort(k, desc).tobytes()
    print(dtype, "desc" if desc else "asc", n, "OK" if good else "MISMATCH", stats(), flush=True)
    ok &= good
# few distinct keys (long runs), constant keys
for mode in ("few", "const"):
    n = (1 << 24) + 99
    k = (rng.integers(0, 3, size=n, dtype=np.uint32) * 0x01010101) if mode == "few" else np.full(n, 0xdeadbeef, np.uint32)
    out = gpu_api.radix_sort(k)
    good = np.array_equal(out, np.sort(k))
    print("u32", mode, n, "OK" if good else "MISMATCH", stats(), flush=True)
    ok &= good
print("ALL_OK" if ok else "FAILED")
