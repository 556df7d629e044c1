"""Where a single stream's GPU time goes (drop-in case): support/sample.xz through one engine with L2 on the device,
(a) whole capture resident, one nrsc5b_process, wall time; (b) the same with CUDA events around every launch."""
import lzma
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nrsc5_b200                                          # noqa: E402
import torch                                               # noqa: E402

cu8 = np.frombuffer(lzma.open(os.path.join(ROOT, "oracle/_ref/sample.xz")).read(), dtype=np.uint8)
cu8 = cu8[: cu8.size & ~3]
dev = torch.from_numpy(cu8.copy()).cuda()
for prof in (False, True):
    with nrsc5_b200.Engine(nstreams=1, input_capacity=cu8.size + 4096, log_capacity=4 << 20) as e:
        e.enable_l2(True)
        best = 1e9
        for rep in range(4):
            e.reset()
            e.set_profiling(prof)
            e.attach_device_input(dev.data_ptr(), cu8.size, cu8.size)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e.process()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        st = e.stats()
        print("profiling" if prof else "plain", "wall ms %.3f" % (best * 1e3), "launches", st.kernel_launches, "blocks", st.blocks,
              e.kernel_times() if prof else "", flush=True)
