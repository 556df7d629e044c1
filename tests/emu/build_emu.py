"""Builds tests/_build/libnrsc5_b200_emu.so: the engine's CUDA sources (nrsc5_b200/csrc/engine.cu, frontend.cu and
the headers they include, unmodified) compiled with g++ against the CPU emulation of the CUDA execution model in
tests/emu/cuda_emu.h.  TEST INFRASTRUCTURE ONLY - the product library is built by nrsc5_b200/build.py with nvcc and
has no CPU path.

The only source transformation is the launch syntax, which is not C++:
    kernel<<<grid, block, smem, stream>>>(args);  ->  emu::launch(dim3(grid), dim3(block), smem, [&]() { kernel(args); });
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "nrsc5_b200", "csrc")
OUTDIR = os.path.join(ROOT, "tests", "_build")
OUT = os.path.join(OUTDIR, "libnrsc5_b200_emu_asan.so" if os.environ.get("EMU_ASAN") else "libnrsc5_b200_emu.so")
UNITS = ["engine.cu", "frontend.cu"]


def _match(s, i, open_ch, close_ch):
    """index of the bracket closing the one at s[i]"""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == open_ch:
            depth += 1
        elif s[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced brackets")


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return [p.strip() for p in parts]


def rewrite_launches(src: str) -> str:
    out, pos = "", 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            return out + src[pos:]
        j = src.index(">>>", i)
        # kernel name: identifier (with optional template arguments) right before <<<
        k = i
        if src[k - 1] == ">":                       # template arguments
            depth = 0
            while True:
                k -= 1
                if src[k] == ">":
                    depth += 1
                elif src[k] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while k > 0 and (src[k - 1].isalnum() or src[k - 1] in "_:"):
            k -= 1
        name = src[k:i]
        cfg = _split_top(src[i + 3:j])
        cfg += ["0"] * (3 - len(cfg)) if len(cfg) < 3 else []
        a0 = src.index("(", j)
        a1 = _match(src, a0, "(", ")")
        args = src[a0 + 1:a1]
        semi = src.index(";", a1)
        out += src[pos:k]
        out += "emu::launch(dim3(%s), dim3(%s), (size_t)(%s), [&]() { %s(%s); })" % (cfg[0], cfg[1], cfg[2], name, args)
        pos = semi


def build(verbose=False):
    os.makedirs(OUTDIR, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in os.listdir(HERE)]
    deps.append(os.path.join(ROOT, "include", "nrsc5_b200.h"))
    if os.path.exists(OUT) and all(os.path.getmtime(p) <= os.path.getmtime(OUT) for p in deps):
        return OUT
    objs = []
    cxx = os.environ.get("CXX", "g++")
    flags = ["-std=c++17", "-O2", "-g", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-w", "-I" + HERE, "-I" + CSRC,
             "-I" + os.path.join(ROOT, "include")]
    if os.environ.get("EMU_ASAN"):          # EMU_ASAN=1: AddressSanitizer build (run python with LD_PRELOAD=libasan.so)
        # + alignment checks: a misaligned short2 / float2 / uint4 access is fatal on the GPU and silent on x86
        flags += ["-fsanitize=address,alignment", "-fno-sanitize-recover=alignment", "-fno-omit-frame-pointer"]
    for u in UNITS:
        gen = os.path.join(OUTDIR, ("asan_" if os.environ.get("EMU_ASAN") else "") + "emu_" + u.replace(".cu", ".cpp"))
        text = open(os.path.join(CSRC, u)).read()
        with open(gen, "w") as f:
            f.write('#line 1 "%s"\n' % os.path.join(CSRC, u))
            f.write(rewrite_launches(text))
        obj = gen.replace(".cpp", ".o")
        cmd = [cxx, *flags, "-c", gen, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        objs.append(obj)
    obj = os.path.join(OUTDIR, ("asan_" if os.environ.get("EMU_ASAN") else "") + "cuda_emu.o")
    subprocess.run([cxx, *flags, "-c", os.path.join(HERE, "cuda_emu.cpp"), "-o", obj], check=True)
    subprocess.run([cxx, "-shared", "-o", OUT, *objs, obj, "-lm"] + (["-fsanitize=address,alignment"] if os.environ.get("EMU_ASAN") else []), check=True)
    return OUT


if __name__ == "__main__":
    print(build(verbose=True))
