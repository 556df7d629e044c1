"""One PROCESS per channel, each pinned to its own physical core, around the unmodified reference
(oracle/_ref/libnrsc5_ref.so) - BASELINE.md §3's CPU baseline.  TEST / BENCH INFRASTRUCTURE ONLY.

Why processes: the reference mallocs and frees 18.7 MB of Viterbi path memory per P1 frame
(reference src/conv_dec.c:440,451); threads of one process serialise on the address-space lock while the kernel maps
and unmaps those pages, separate processes do not.
"""
import multiprocessing as mp
import os
import time

import numpy as np

import reftap


def physical_cores():
    """One logical CPU per physical core, among those this process may run on."""
    allowed = sorted(os.sched_getaffinity(0))
    seen, out = set(), []
    try:
        topo = {}
        cpu = None
        phys = core = None
        for line in open("/proc/cpuinfo"):
            if line.startswith("processor"):
                cpu = int(line.split(":")[1])
                phys = core = None
            elif line.startswith("physical id"):
                phys = int(line.split(":")[1])
            elif line.startswith("core id"):
                core = int(line.split(":")[1])
                topo[cpu] = (phys, core)
        for c in allowed:
            key = topo.get(c, (None, c))
            if key not in seen:
                seen.add(key)
                out.append(c)
    except OSError:
        out = allowed
    return out or allowed


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def fft_backend():
    return "oracle/shim/fftshim.c (radix-2; FFTW 3.3.10 is not installed in this image)"


def _child(buf, cpu, mode, reps, barrier, q):
    try:
        os.sched_setaffinity(0, {cpu})
    except OSError:
        pass
    reftap.lib()
    barrier.wait()
    t0 = time.monotonic()
    reftap.bench([buf], mode=mode, reps=reps)
    q.put((t0, time.monotonic()))


def bench_processes(bufs, mode=reftap.MODE_FM, reps=1, cpus=None):
    """Decode bufs[i] `reps` times in process i (pinned to cpus[i]); returns wall seconds from the common start to the
    last process's end."""
    cpus = cpus or physical_cores()
    n = min(len(bufs), len(cpus))
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(n)
    q = ctx.Queue()
    procs = [ctx.Process(target=_child, args=(np.ascontiguousarray(bufs[i]), cpus[i], mode, reps, barrier, q)) for i in range(n)]
    for p in procs:
        p.start()
    spans = [q.get() for _ in range(n)]
    for p in procs:
        p.join()
    return max(t1 for _, t1 in spans) - min(t0 for t0, _ in spans), n
