"""How much would overlapping decode groups with the next front-end pass buy?  Proxy: the headline's 128 streams as
ONE engine (passes strictly serial) against TWO engines of 64 streams on two CUDA streams driven by two host threads
(one engine's decode groups run beside the other's front end).  Timing only."""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                               # noqa: E402
import nrsc5_b200                                          # noqa: E402
import torch                                               # noqa: E402

caps = bench.make_captures(4, 4)
views, nbytes = bench.stream_views(caps, 128, 0)
import numpy as np                                         # noqa: E402
host = np.stack([np.ascontiguousarray(v) for v in views])
dev = torch.from_numpy(host).cuda()


def run(groups, reps=8):
    engines = []
    for g in range(groups):
        S = 128 // groups
        e = nrsc5_b200.Engine(nstreams=S, input_capacity=nbytes + 4096, log_capacity=1 << 20)
        st = torch.cuda.Stream()
        e.set_cuda_stream(st.cuda_stream)
        engines.append((e, dev[g * S:(g + 1) * S], st))

    def one(e, d, st):
        e.attach_device_input(d.data_ptr(), nbytes, nbytes)
        e.rewind()
        e.process()

    best = 1e9
    for r in range(reps + 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=one, args=x) for x in engines]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        if r >= 2:
            best = min(best, time.perf_counter() - t0)
    for e, _, _ in engines:
        e.close()
    return best * 1e3


for g in (1, 2, 4):
    print("engines", g, "x", 128 // g, "streams: ms per step %.3f" % run(g), flush=True)
