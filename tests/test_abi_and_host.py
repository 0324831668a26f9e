"""CPU-only checks of the boundary: the shared library loads, exports every
symbol include/flagstats_cuda.h declares, refuses to compute without a device
(no CPU fallback), and the Python host mirror behaves like pyflagstats."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fs():
    from libflagstats_b200 import build
    build.build()
    import libflagstats_b200 as m
    return m


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "flagstats_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b((?:FLAGSTAT|POSPOPCNT)_cuda\w*)\s*\(", src)
    return sorted(set(n for n in names if n != "FLAGSTATS_cuda_func"))


def test_library_exports_every_declared_symbol(fs):
    from libflagstats_b200 import _capi
    names = declared_symbols()
    assert len(names) >= 30
    handle = C.CDLL(_capi.SO_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/flagstats_cuda.h but not exported"
    assert sorted(_capi.SIGNATURES) == names, "ctypes table and header drifted apart"
    assert b"sm_100a" in fs.lib().FLAGSTAT_cuda_version()


def test_no_cpu_fallback_without_a_device(fs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert fs.available() == 0
    a = np.arange(4096, dtype=np.uint16)
    with pytest.raises(fs.FlagstatCudaError) as ei:
        fs.flagstat_u64(a)
    assert ei.value.code == -1
    f = np.full(32, 5, np.uint32)
    assert fs.lib().FLAGSTAT_cuda(a.ctypes.data, a.size, f.ctypes.data_as(C.POINTER(C.c_uint32))) == -1
    assert (f == 5).all()  # flags untouched on failure
    assert fs.lib().FLAGSTAT_cuda_strerror(-1) == b"no usable CUDA device"
    with pytest.raises(fs.FlagstatCudaError):
        fs.BlockStream(0)


def test_python_surface_mirrors_pyflagstats_errors(fs):
    # python/libflagstats.pyx:9-13: type and dtype are checked before anything else
    with pytest.raises(ValueError, match="numpy.ndarray"):
        fs.flagstats([1, 2, 3])
    with pytest.raises(ValueError, match="uint16"):
        fs.flagstats(np.arange(10, dtype=np.int16))


def test_result_dict_schema_matches_pyx(fs, golden):
    e = golden["kat_e"]
    f = np.array(e["cuda_expected"], np.uint64)
    d = fs.counters_to_dict(f, e["spec"]["n"])
    assert list(d) == ["n_values", "passed", "failed"]
    assert list(d["passed"])[:15] == fs.SAM_FLAG_NAMES and list(d["failed"]) == fs.SAM_FLAG_NAMES
    assert int(d["passed"]["mapped"]) == e["readme"]["mapped"]
    assert int(d["passed"]["paired_in_seq"]) == e["readme"]["paired"]
    rep = fs.samtools_report(f).splitlines()
    # README.md:179-189, the FLAG-derivable lines
    assert rep[0] == "824541892 + 0 in total (QC-passed reads + QC-failed reads)"
    assert rep[1] == "0 + 0 secondary" and rep[2] == "5393628 + 0 supplementary"
    assert rep[3] == "0 + 0 duplicates" and rep[4] == "805383403 + 0 mapped (97.68% : N/A)"
    assert rep[5] == "819148264 + 0 paired in sequencing"
    assert rep[6] == "409574132 + 0 read1" and rep[7] == "409574132 + 0 read2"
    assert rep[8] == "781085884 + 0 properly paired (95.35% : N/A)"
    assert rep[9] == "797950890 + 0 with itself and mate mapped"
    assert rep[10] == "2038885 + 0 singletons (0.25% : N/A)"


def test_shard_ranges_tile_the_column():
    from libflagstats_b200.sharded import shard_range
    for n in (0, 1, 7, 8, 9, 1000, 824541892, 2 ** 34, 2 ** 34 + 5):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and a <= b
            assert all(lo % 8 == 0 for lo, _ in edges)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_kat_e_shards_in_golden_use_the_same_ranges(golden):
    from libflagstats_b200.sharded import shard_range
    n = golden["kat_e"]["spec"]["n"]
    for r, s in enumerate(golden["kat_e"]["shards8"]):
        lo, hi = shard_range(n, 8, r)
        assert (lo, hi - lo) == (s["start"], s["n"])


def test_product_never_touches_the_oracle():
    """oracle/ is the checker: only tests/, __graft_entry__.smoke()/build() and bench.py's CPU legs
    may import, link or execute it.  Static check over the product sources and the build recipe."""
    import re
    pkg = os.path.join(ROOT, "libflagstats_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle/|liboracle|oracle_[a-z]+\(", re.M)
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                continue
            text = open(os.path.join(dirpath, f), encoding="utf-8").read()
            # comments may NAME the oracle (e.g. "equals the CPU oracle byte for byte"); code may not use it
            code = "\n".join(ln for ln in text.splitlines()
                             if not ln.lstrip().startswith(("//", "#", "*", "/*", '"""')) and "oracle/flagstat_oracle.c)" not in ln)
            assert not pat.search(code), os.path.join(dirpath, f)
    # the library links nothing from oracle/ either
    from libflagstats_b200 import build as b
    assert not any("oracle" in x for x in b.SOURCES + b.HEADERS + b.NVCC_FLAGS)
    # bench.py: the oracle is imported inside the cpu_* functions only (cpu_baseline legs / --impl
    # reference); the GPU legs build their inputs without it (tools/containers.py)
    bench = open(os.path.join(ROOT, "bench.py"), encoding="utf-8").read()
    uses = [m.start() for m in re.finditer(r"from oracle import|import oracle", bench)]
    assert uses
    for u in uses:
        d = bench.rfind("\ndef ", 0, u)
        assert bench[d + 1:].startswith("def cpu_"), bench[d + 1:d + 60]
    containers = open(os.path.join(ROOT, "tools", "containers.py"), encoding="utf-8").read()
    assert not re.search(r"^\s*(from|import)\s+oracle\b", containers, re.M)


def test_header_is_plain_c99_and_cxx11(tmp_path):
    """include/flagstats_cuda.h is what a C host binds (the reference is a C header): it must
    compile on its own as strict C99 / C11 and C++11, no CUDA or torch types, no extensions."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None or shutil.which("g++") is None:
        pytest.skip("no gcc / g++")
    src = '#include "flagstats_cuda.h"\nint main(void) { return sizeof(FLAGSTAT_cuda_bam_flagstat) == 26 * sizeof(long long) ? 0 : 1; }\n'
    inc = os.path.join(ROOT, "include")
    for cc, std, ext in (("gcc", "c99", "c"), ("gcc", "c11", "c"), ("g++", "c++11", "cpp")):
        f = tmp_path / f"h_{std.replace('+', 'x')}.{ext}"
        f.write_text(src)
        r = subprocess.run([cc, f"-std={std}", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", inc, "-c", str(f),
                            "-o", str(tmp_path / "h.o")], capture_output=True, text=True)
        assert r.returncode == 0, (cc, std, r.stderr)
    import re
    text = open(os.path.join(inc, "flagstats_cuda.h"), encoding="utf-8").read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # comments may say "a cudaStream_t travels as void*"
    for banned in ("cuda_runtime", "cudaStream_t", "cudaError_t", "torch", "#include <cuda"):
        assert banned not in code, banned


def test_tools_and_session_scripts_parse():
    """tools/ holds the measurement scripts the GPU sessions run (there is no second chance on a
    box that is charged by the minute): every Python tool compiles, every shell script
    passes `bash -n`."""
    import glob
    import subprocess
    for f in glob.glob(os.path.join(ROOT, "tools", "*.py")) + [os.path.join(ROOT, "bench.py"),
                                                                 os.path.join(ROOT, "__graft_entry__.py")]:
        compile(open(f, encoding="utf-8").read(), f, "exec")
    for f in glob.glob(os.path.join(ROOT, "tools", "*.sh")) + glob.glob(os.path.join(ROOT, "integration", "*.sh")):
        r = subprocess.run(["bash", "-n", f], capture_output=True, text=True)
        assert r.returncode == 0, (f, r.stderr)
