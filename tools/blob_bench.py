"""f3 microbench: DataBlob assembly for 4 MiB chunks of an 8 GiB device-resident image, half of it zero runs, into PINNED host
memory through the raw C calls (no Python-side allocation or slicing inside the timed region)."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pbs_plus_b200 as pg  # noqa: E402

e = pg.Engine(0)
L_ = e._L
n = 8 << 30
d = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
d[1 << 30: 5 << 30] = 0
CH = 4 << 20
off = np.arange(0, n, CH, dtype=np.uint64)
ln = np.full(len(off), CH, dtype=np.uint64)
boff = (np.arange(len(off), dtype=np.uint64) * np.uint64(CH + 12))
out = e.host_alloc(int(len(off)) * (CH + 12))
blen = np.zeros(len(off), dtype=np.uint64)
crc = np.zeros(len(off), dtype=np.uint32)
for name in ("crc32_batch", "blob_encode_batch", "blob_encode_batch_z"):
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if name == "crc32_batch":
            rc = L_.pbsgpu_crc32_batch(e._h, d.data_ptr(), off.ctypes.data, ln.ctypes.data, len(off), crc.ctypes.data)
            size = 0
        elif name == "blob_encode_batch":
            rc = L_.pbsgpu_blob_encode_batch(e._h, d.data_ptr(), off.ctypes.data, ln.ctypes.data, len(off), out.ctypes.data,
                                             boff.ctypes.data, crc.ctypes.data)
            size = len(off) * (CH + 12)
        else:
            rc = L_.pbsgpu_blob_encode_batch_z(e._h, d.data_ptr(), off.ctypes.data, ln.ctypes.data, len(off), out.ctypes.data,
                                               boff.ctypes.data, blen.ctypes.data, crc.ctypes.data)
            size = int(blen.sum())
        dt = time.perf_counter() - t0
        assert rc == 0, rc
    print(f"{name}: {n / dt / 1e9:.1f} GB/s of input ({dt * 1e3:.0f} ms, third run), {size / 2**30:.2f} GiB written to pinned host memory")
e.host_free(out)
e.close()
