"""Build tests/emu/_build/libmaed_emu.so: the CUDA-core translation units and the host orchestration of maed_b200/csrc
compiled by g++ against the CUDA-on-CPU shim (cuda_emu.h), plus CPU restatements of the tensor-core kernels' contracts
(tc_stubs.cpp).  TEST INFRASTRUCTURE ONLY: the product library is maed_b200/libmaed_b200.so (nvcc, sm_100a); nothing under
maed_b200/ imports this module or loads the emulator library.

The sources are used as they are; two CUDA-only spellings are rewritten on the way into _build/:
  kernel<<<grid, block, smem, stream>>>(args)   ->  ::emu::Launch(grid, block, smem, stream).call(kernel, args)
  extern __shared__ T name[];                    ->  T* name = reinterpret_cast<T*>(::emu::t_dyn_smem);
"""
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "maed_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libmaed_emu.so")
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"

# compiled from the real sources through the shim
CU_SOURCES = ["kernels.cu", "decoder.cu", "bwd_kernels.cu", "bwd_kernels2.cu", "attention_bwd.cu", "smpl.cu", "loss.cu", "decode_bwd.cu",
              "cnn_kernels.cu", "cnn_engine.cu", "engine.cu", "train.cu", "capi.cu"]
# tensor-core / TMA translation units replaced by tc_stubs.cpp
REPLACED = ["gemm_host.cu", "gemm_gn_sm100.cu", "stem_sm100.cu", "attention.cu", "gemm_splitk_sm100.cu"]
EMU_SOURCES = ["cuda_emu.cpp", "tc_stubs.cpp"]

_LAUNCH = re.compile(r"([A-Za-z_]\w*(?:<[^;<>(){}]*>)?)\s*<<<(.*?)>>>\s*\(", re.S)
_EXTERN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\s*\d+\s*\)\s+)?([A-Za-z_][\w:]*)\s+(\w+)\s*\[\s*\]\s*;")


def translate(text):
    def launch(m):
        return "::emu::Launch(%s).call(%s, " % (m.group(2), m.group(1))
    text = _LAUNCH.sub(launch, text)
    text = re.sub(r"\.call\(([^()]*?),\s*\)", r".call(\1)", text)        # kernels without parameters
    text = _EXTERN_SMEM.sub(lambda m: "%s* %s = reinterpret_cast<%s*>(::emu::t_dyn_smem);" % (m.group(1), m.group(2), m.group(1)),
                            text)
    assert "<<<" not in text, "untranslated kernel launch"
    # gemm_sm100.cuh is PTX; the host code only needs its OUT_* / ACT_* enums (extracted into _build/emu_gemm_enums.h)
    text = text.replace('#include "gemm_sm100.cuh"', '#include "emu_gemm_enums.h"')
    return text


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("emulator build failed")


def build(force=False, verbose=False):
    os.makedirs(OUT, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(HERE, f) for f in sorted(os.listdir(HERE))
                                                                         if f.endswith((".h", ".cpp", ".py"))]
    deps.append(os.path.join(ROOT, "include", "maed_b200.h"))
    h = hashlib.sha1((os.environ.get("MAED_EMU_ASAN", "") + "|" + os.environ.get("MAED_EMU_UBSAN", "")).encode())
    try:                                    # -march=native: never reuse a library built on another CPU model
        with open("/proc/cpuinfo") as f:
            h.update("".join(l for l in f if l.startswith(("model name", "flags"))).encode())
    except OSError:
        pass
    for d in deps:
        with open(d, "rb") as f:
            h.update(f.read())
    stamp = os.path.join(OUT, "stamp")
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return LIB
    flags = ["-O2", "-g1", "-std=c++17", "-fPIC", "-march=native", "-fno-strict-aliasing", "-Wno-attributes",
             "-Wno-unknown-pragmas", "-Wno-deprecated-declarations",
             "-include", os.path.join(HERE, "cuda_emu.h"), "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include"),
             "-I", CUDA_INC]
    with open(os.path.join(CSRC, "gemm_sm100.cuh")) as f:
        enums = re.findall(r"^enum : int \{[^}]*\};$", f.read(), re.M)
    assert len(enums) == 2, enums
    with open(os.path.join(OUT, "emu_gemm_enums.h"), "w") as f:
        f.write("// generated from maed_b200/csrc/gemm_sm100.cuh by build_emu.py\n#pragma once\nnamespace maed {\n%s\n}\n" % "\n".join(enums))
    flags += ["-I", OUT]
    if os.environ.get("MAED_EMU_ASAN"):      # one-off memory checking (tests/emu/README.md): heap redzones around every tensor
        flags += ["-fsanitize=address", "-fno-omit-frame-pointer", "--param", "asan-stack=0"]
    if os.environ.get("MAED_EMU_UBSAN"):     # misaligned vector accesses (float4 / uint4 / uint2 ...) fault on the GPU but not on x86:
        flags += ["-fsanitize=alignment", "-fno-sanitize-recover=alignment"]   # make them abort here as well
    objs, jobs = [], []
    for name in CU_SOURCES:
        with open(os.path.join(CSRC, name)) as f:
            src = '#line 1 "%s"\n' % os.path.join(CSRC, name) + translate(f.read())
        cpp = os.path.join(OUT, name.replace(".cu", ".emu.cpp"))
        with open(cpp, "w") as f:
            f.write(src)
        jobs.append((cpp, cpp[:-4] + ".o"))
    for name in EMU_SOURCES:
        jobs.append((os.path.join(HERE, name), os.path.join(OUT, name[:-4] + ".o")))
    procs = []
    for src, obj in jobs:
        cmd = ["g++"] + flags + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for cmd, p in procs:
        out = p.communicate()[0]
        if p.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + out[-6000:] + "\n")
            failed = True
        elif verbose and out.strip():
            print(out[-3000:])
    if failed:
        raise RuntimeError("emulator build failed")
    _run(["g++", "-shared", "-o", LIB] + objs + ["-lpthread", "-Wl,-Bsymbolic"] +
         (["-fsanitize=address"] if os.environ.get("MAED_EMU_ASAN") else []) +
         (["-fsanitize=alignment"] if os.environ.get("MAED_EMU_UBSAN") else []))
    with open(stamp, "w") as f:
        f.write(h.hexdigest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
