"""TEST INFRASTRUCTURE ONLY — builds `tests/host_emul/_build/libbtcdet_b200_emul.so`: the C-ABI entry points and kernels of
the index / pooling / RoI source files (btcdet_b200/csrc/*.cu, unchanged on disk) compiled for the HOST with g++ under the
lock-step emulation of cuda_emul.h, so that the CPU test-suite can drive the real `btc_*` entry points with numpy buffers
standing in for device memory and compare them with the oracle / the golden vectors where no GPU exists.

What this script does to a source file (in a temporary copy only): every kernel launch
`name<<<grid, block, smem, stream>>>(args);` becomes `emul::launch(grid, block, smem, [&] { name(args); });` and
`extern __shared__ T name[];` becomes a pointer to the per-launch dynamic shared buffer.  Nothing else is touched.
The tcgen05 files (inline PTX) are not part of it.  Never shipped, never imported by btcdet_b200/ or spconv/.
"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "btcdet_b200", "csrc")
OUT = os.path.join(HERE, "_build")
SAN = os.environ.get("BTC_EMUL_SANITIZE", "")      # e.g. "address": an AddressSanitizer build (run pytest under LD_PRELOAD=libasan)
LIB = os.path.join(OUT, "libbtcdet_b200_emul%s%s.so" % ("_" + SAN if SAN else "", "_fibers" if SAN and os.environ.get("BTC_EMUL_FIBERS") else ""))
FILES = ["coord_index.cu", "voxelize.cu", "rulebook.cu", "pool_dense.cu", "points_transform.cu", "roi_pool.cu",
         "sparse_conv.cu", "iou3d_nms.cu", "occ_masks.cu", "box_masks.cu"]


def _match(text, i, open_ch, close_ch):
    """index just past the bracket that closes the one at text[i]"""
    depth = 0
    for j in range(i, len(text)):
        if text[j] == open_ch:
            depth += 1
        elif text[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j + 1
    raise ValueError("unbalanced %s%s" % (open_ch, close_ch))


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def transform(text):
    out, pos = "", 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            break
        # kernel name (with optional template arguments) right before <<<
        j = i
        while text[j - 1].isspace():
            j -= 1
        if text[j - 1] == ">":
            depth, k = 0, j - 1
            while True:
                if text[k] == ">":
                    depth += 1
                elif text[k] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                k -= 1
            j = k
        k = j
        while text[k - 1].isalnum() or text[k - 1] in "_:":
            k -= 1
        name = text[k:i].strip()
        e = text.index(">>>", i)
        cfg = _split_top(text[i + 3:e])
        while len(cfg) < 3:
            cfg.append("0")
        a0 = text.index("(", e)
        a1 = _match(text, a0, "(", ")")
        semi = text.index(";", a1)
        out += text[pos:k] + "emul::launch(dim3(%s), dim3(%s), (size_t)(%s), [&] { %s%s; })" % (
            cfg[0], cfg[1], cfg[2], name, text[a0:a1])
        pos = semi
    out += text[pos:]
    out = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w\s]+?)\s+(\w+)\s*\[\s*\]\s*;",
                 r"\1* \2 = (\1*)emul::dyn_smem;", out)
    return out


STUBS = r'''
// host definitions of the few CUDA runtime calls the entry points make ("device memory" is host memory here)
extern "C" {
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
// the tcgen05 tile is not part of the emulated build: the shim's shape query says "not supported" and the fp32 FFMA tile runs
int btc_sparse_conv_tc_supported(int, int, int) { return 0; }
int btc_sparse_conv_tc_split_supported(int, int, int, int, int) { return 0; }
}
'''


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "cuda_emul.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    objs = []
    for f in FILES:
        src = open(os.path.join(CSRC, f)).read()
        cpp = os.path.join(OUT, f.replace(".cu", ".emul.cpp"))
        with open(cpp, "w") as fh:
            fh.write('#include "cuda_emul.h"\n' + (STUBS if f == FILES[0] else "") + transform(src))
        obj = cpp.replace(".cpp", (".%s%s.o" % (SAN, "f" if os.environ.get("BTC_EMUL_FIBERS") else "")) if SAN else ".o")
        res = subprocess.run(["g++", "-std=c++20", "-O1", "-fPIC", "-pthread", "-ffp-contract=off", "-w", "-DBTC_SM=100"] +
                             (["-g", "-fno-omit-frame-pointer", "-fsanitize=" + SAN] + ([] if os.environ.get("BTC_EMUL_FIBERS") else ["-DEMUL_THREADS"])
                              if SAN else []) + [
                              "-I" + HERE, "-I" + CSRC, "-I/usr/local/cuda/include", "-c", cpp, "-o", obj],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("g++ failed for %s:\n%s" % (f, res.stderr[:4000]))
        objs.append(obj)
    subprocess.run(["g++", "-shared", "-pthread", "-Wl,-Bsymbolic"] + (["-fsanitize=" + SAN] if SAN else []) + ["-o", LIB] + objs,
                   check=True, capture_output=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
