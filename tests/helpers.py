"""Shared test helpers: regenerate golden inputs from their generator spec."""
import numpy as np

from oracle import oracle as O

CORE19 = list(O.CORE19)
CORE20 = list(O.CORE20)
UNTOUCHED = [i for i in range(32) if i not in CORE20]


def mt19937_inmemory(n):
    """benchmark/inmemory.cpp:108-116 input stream (mt19937 seed 0, >> 20)."""
    bg = np.random.MT19937()
    bg._legacy_seeding(0)
    return (bg.random_raw(n) >> 20).astype(np.uint16)


def make_input(spec):
    g = spec["gen"]
    if g == "arange":
        return (np.arange(spec["n"], dtype=np.uint32) & 0xFFFF).astype(np.uint16)
    if g == "mt19937_inmemory":
        return mt19937_inmemory(spec["n"])
    if g == "uniform":
        return O.synth_uniform(spec.get("start", 0), spec["n"], spec["seed"], spec["mask"])
    if g == "hiseqx":
        return O.synth_hiseqx(spec.get("start", 0), spec["n"], spec.get("seed", 0),
                              spec.get("qcfail_ppm", 0))
    if g == "const":
        return np.full(spec["n"], spec["value"], np.uint16)
    raise ValueError(g)


def offset_copy(a, off):
    """Copy of `a` whose base address is `off` elements past a 256-byte boundary."""
    raw = np.empty(a.size + 128 + off, np.uint16)
    base = (-(raw.ctypes.data // 2)) % 128
    view = raw[base + off: base + off + a.size]
    view[:] = a
    assert a.size == 0 or (view.ctypes.data - 2 * off) % 256 == 0
    return view
