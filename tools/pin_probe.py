"""How fast is cudaHostRegister on this box, and what do registered buffers buy (development aid)?
   python tools/pin_probe.py [GB]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from dampr_b200 import device as dev
from dampr_b200 import runner as runner_mod


def main():
    gb = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
    n = int(gb * (1 << 30)) // 16
    ctx = runner_mod.get_ctx()
    kv = ctx.synth_kv(3, n, n)
    ctx.sync()
    out = {}
    buf = np.empty((n, 2), dtype=np.uint64)
    t0 = time.perf_counter(); kv.records_into(buf); out["d2h_pageable_cold_GBps"] = buf.nbytes / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter(); kv.records_into(buf); out["d2h_pageable_warm_GBps"] = buf.nbytes / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter(); kv.upload(0, buf, n); ctx.sync(); out["h2d_pageable_GBps"] = buf.nbytes / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter(); ok = dev.host_register(buf); dt = time.perf_counter() - t0
    out["register_ok"] = ok
    out["register_touched_GBps"] = buf.nbytes / dt / 1e9
    t0 = time.perf_counter(); kv.records_into(buf); out["d2h_registered_GBps"] = buf.nbytes / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter(); kv.upload(0, buf, n); ctx.sync(); out["h2d_registered_GBps"] = buf.nbytes / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter(); dev.host_unregister(buf); out["unregister_GBps"] = buf.nbytes / (time.perf_counter() - t0) / 1e9
    fresh = np.empty((n, 2), dtype=np.uint64)
    t0 = time.perf_counter(); ok = dev.host_register(fresh); dt = time.perf_counter() - t0
    out["register_untouched_GBps"] = fresh.nbytes / dt / 1e9
    dev.host_unregister(fresh)
    # registration in 1 GiB pieces next to a running transfer
    import threading
    fresh2 = np.empty((n, 2), dtype=np.uint64)
    fresh2[:] = 0
    flat = fresh2.reshape(-1)
    step = (1 << 30) // 8
    def reg():
        for lo in range(0, len(flat), step):
            dev.host_register(flat[lo:lo + step])
    th = threading.Thread(target=reg)
    t0 = time.perf_counter(); th.start(); kv.records_into(buf); t1 = time.perf_counter(); th.join(); t2 = time.perf_counter()
    out["d2h_pageable_while_registering_GBps"] = buf.nbytes / (t1 - t0) / 1e9
    out["register_chunks_while_copying_GBps"] = fresh2.nbytes / (t2 - t0) / 1e9
    for lo in range(0, len(flat), step):
        dev.host_unregister(flat[lo:lo + step])
    print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
