"""CPU tests of the drop-in boundary: the shared library loads without a GPU and exports every symbol that
include/nt_b200.h declares, plus the reference's mangled C++ launcher names (src/cuda/kernels.h:10-74)."""
import re
import subprocess
from pathlib import Path

from ntransformer_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "nt_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(nt_b200_\w+|nt_cuda_\w+)\s*\(", text))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.lib()
    decl = declared_symbols()
    assert len(decl) >= 30
    for name in sorted(decl):
        assert hasattr(lib, name), f"{name} declared in include/nt_b200.h but not exported"
    assert decl == set(_lib.SIGNATURES), "python binding table and header disagree"
    assert lib.nt_b200_version().startswith(b"ntransformer_b200")
    assert lib.nt_b200_xq_bytes(8192) == 3 * 8192 + 8192 // 32 * 4 + 8192 // 16 * 4


def test_reference_cxx_launcher_names_are_exported():
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    for fn in _lib.CXX_LAUNCHERS:
        assert re.search(rf"\bnt::cuda::{fn}\(", out), f"nt::cuda::{fn} missing"
    # the exact signature the reference host code calls (kernels.h:34-36)
    assert "nt::cuda::launch_gemv(float*, void const*, float const*, int, int, nt::DType, void*)" in out


def test_sass_is_sm100a_with_tma_and_dp4a():
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass, "TMA bulk copies missing from the GEMV"
    assert "IDP.4A" in sass
    assert "SYNCS.ARRIVE.TRANS64" in sass
    # prefill: tcgen05 MMA fed by tensor-map TMA with TMEM loads in the epilogue; tiled attention on the warp-level MMA
    for mnemonic in ("UTCHMMA", "UTMALDG.2D", "LDTM", "UTCBAR", "HMMA.16816.F32", "LDSM.16.MT88.4"):
        assert mnemonic in sass, f"{mnemonic} missing"


def test_persistent_decode_kernel_sass():
    """decode_step_kernel (csrc/engine/decode_megakernel.cu): TMA bulk copies + dp4a in the GEMV phases, GPU-scope release/acquire
    for the grid barrier, strong L2 loads for data written by other CTAs, system-scope stores/fences for the peer exchange."""
    obj = ROOT / "ntransformer_b200" / "_build" / "engine__decode_megakernel.cu.o"
    sass = subprocess.run(["cuobjdump", "-sass", str(obj if obj.exists() else _lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    body = sass[sass.index("decode_step_kernel"):]
    nxt = body.find("Function :", 10)
    body = body if nxt < 0 else body[:nxt]
    for mnemonic in ("UBLKCP", "IDP.4A", "SYNCS.ARRIVE.TRANS64", "REDG.E.ADD.STRONG.GPU", "LDG.E.STRONG.GPU", "CCTL.IVALL", "MEMBAR.SC.SYS",
                     "STG.E.STRONG.SYS", "LDG.E.STRONG.SYS", "FENCE.VIEW.ASYNC"):
        assert mnemonic in body, f"{mnemonic} missing from decode_step_kernel"
    for forbidden in ("HMMA", "UTCHMMA", "IMMA"):                     # no tensor cores on the decode path (north_star)
        assert forbidden not in body, forbidden


def test_emulator_hooks_stay_out_of_the_product():
    """tests/cusim compiles some product sources with -DNT_CUSIM; the product build never defines it and carries no emulator code."""
    from ntransformer_b200 import build as nb

    assert not any("NT_CUSIM" in f for f in nb.COMMON + nb.ARCH)
    syms = subprocess.run(["nm", "-D", "-C", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    assert "cusim" not in syms
    pkg = ROOT / "ntransformer_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.cuh")):
        if "_build" in f.parts:
            continue
        text = f.read_text(errors="replace")
        if re.search(r'cusim::|#include "cusim', text):               # code (not a comment pointing at tests/cusim)
            assert "#ifdef NT_CUSIM" in text, f


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under ntransformer_b200/ may import, link or execute it."""
    pkg = ROOT / "ntransformer_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.cuh")):
        if "_build" in f.parts:
            continue
        text = f.read_text(errors="replace")
        assert "nt_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f
    out = subprocess.run(["ldd", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_quarter_block_gemv_and_peer_exchange_sass():
    """gemv_kqq_kernel (csrc/gemv_kquant_q.cu): TMA bulk copies + dp4a, the tensor-parallel epilogue's 64-bit system-scope stores
    {sequence : value} straight into peer memory, no tensor-core ops; xchg_reduce_kernel polls the same words with 64-bit
    system-scope loads (csrc/engine/peer_xchg.cu)."""
    b = ROOT / "ntransformer_b200" / "_build"
    q = subprocess.run(["cuobjdump", "-sass", str(b / "gemv_kquant_q.cu.o")], capture_output=True, text=True, check=True).stdout
    for mnemonic in ("UBLKCP", "IDP.4A", "SYNCS.ARRIVE.TRANS64", "STG.E.64.STRONG.SYS"):
        assert mnemonic in q, f"{mnemonic} missing from the quarter-block GEMV"
    for forbidden in ("HMMA", "UTCHMMA", "IMMA"):
        assert forbidden not in q, forbidden
    x = subprocess.run(["cuobjdump", "-sass", str(b / "engine__peer_xchg.cu.o")], capture_output=True, text=True, check=True).stdout
    assert "LDG.E.64.STRONG.SYS" in x
