"""The default kernel builds its keep-mask with fp16 arithmetic on the FMA pipe
(libflagstats_b200/csrc/flagstat_kernels.cuh: mask_select_f, fail_gate_f).  This is the
exactness proof of that arithmetic, on the CPU: a numpy model of the six instructions
(round-to-nearest-even fp16 results from float64 intermediates = one rounding per
fma / mul, subnormals kept) over ALL 65,536 values of a FLAG word, against the
reference rule (libflagstats.h:118-142) written as a keep-mask.  The two records of a
packed register are independent lanes of every f16x2 instruction, so one half suffices.
The GPU side of the same statement is tests/test_cuda_parity.py::
test_every_record_value_in_both_register_halves.
"""
import numpy as np


def _f(bits):
    return float(np.array(bits, dtype=np.uint16).view(np.float16))


def _rn16(x):
    return x.astype(np.float16)


def _fma(a, b, c, sat=False):
    r = _rn16(a * b + c).astype(np.float64)
    if sat:
        r = _rn16(np.clip(r, 0.0, 1.0)).astype(np.float64)
    return r


def test_keep_mask_and_fail_gate_are_exact_for_every_flag_word():
    w = np.arange(65536, dtype=np.uint32).astype(np.uint16)
    q = w & 0x0905
    qf = q.view(np.float16).astype(np.float64)
    g1 = (q == 0x0001).astype(np.float64)          # set.eq.f16x2.f16x2 -> 1.0 / 0.0
    x1 = (q == 0x0005).astype(np.float64)
    z = _fma(qf, _f(0x7A00), _f(0xC608), sat=True)  # 1.0 iff SECONDARY and SUPPLEMENTARY
    assert set(z.tolist()) == {0.0, 1.0}
    assert ((z == 1.0) == ((w & 0x0900) == 0x0900)).all()
    k = _fma(x1, _f(0x0300), _f(0x0F04))
    k = _fma(g1, _f(0x032C), k)
    k = _rn16(z * _f(0x8D43) + k).view(np.uint16)
    assert sorted(set(k.tolist())) == [0x0704, 0x0F04, 0x0FC4, 0x0FCF]
    y = w & k

    paired, unmap = (w & 1) != 0, (w & 4) != 0
    sec, supp = (w & 0x100) != 0, (w & 0x800) != 0
    kk = paired & ~sec & ~supp                      # the third branch of the reference's else-if chain
    keep = np.full(65536, 0x0F04, dtype=np.uint16)  # UNMAP, SEC, QCFAIL, DUP, SUPP: counted for every record
    keep[kk] |= 0x00C0                              # READ1 / READ2
    keep[kk & ~unmap] |= 0x000B                     # pair-map marker, PROPER_PAIR, MUNMAP
    keep[sec & supp] &= 0xF7FF                      # `else if`: SUPPLEMENTARY only when not SECONDARY
    assert (y == (w & keep)).all()
    assert ((y & 0xF030) == 0).all()                # clean positions 4, 5, 12..15: finite, non-negative fp16

    f1 = _rn16((w & 0x0200).view(np.float16).astype(np.float64) * _f(0x7800)).astype(np.float64)
    assert set(f1.tolist()) == {0.0, 1.0}
    yf = _rn16(y.view(np.float16).astype(np.float64) * f1).view(np.uint16)
    assert (yf == np.where((w & 0x0200) != 0, y, 0)).all()


def test_samtools_form_counts_n_pair_all_at_position_4_for_every_flag_word():
    """mask_select_fx (kSamtools): the class constants gain bit 4 and the final LOP3 is
    k & (w | 0x0010), so position 4 is 1 exactly for the records flagstat_loop counts in
    n_pair_all (benchmark/flagstats.cpp:54-59) and every other position is mask_select_f's."""
    w = np.arange(65536, dtype=np.uint32).astype(np.uint16)
    q = w & 0x0905
    qf = q.view(np.float16).astype(np.float64)
    g1 = (q == 0x0001).astype(np.float64)
    x1 = (q == 0x0005).astype(np.float64)
    paired = (w & 1) != 0
    sec, supp = (w & 0x100) != 0, (w & 0x800) != 0
    kk = paired & ~sec & ~supp
    for has_sec in (True, False):
        k = _fma(x1, _f(0x0340), _f(0x0F04))
        k = _fma(g1, _f(0x036C), k)
        if has_sec:
            z = _fma(qf, _f(0x7A00), _f(0xC608), sat=True)
            k = _rn16(z * _f(0x8D43) + k)
        else:
            k = _rn16(k)
        k = k.view(np.uint16)
        assert sorted(set(k.tolist())) == ([0x0704] if has_sec else []) + [0x0F04, 0x0FD4, 0x0FDF]
        y = k & (w | 0x0010)
        sel = np.ones(65536, bool) if has_sec else ~sec   # the no-SECONDARY form only sees such batches
        assert (((y & 0x0010) != 0) == kk)[sel].all()
        assert ((y & 0xF020) == 0).all()                 # positions 5, 12..15 stay clean for fail_gate_f

        # all other positions: the plain keep-mask
        unmap = (w & 4) != 0
        keep = np.full(65536, 0x0F04, dtype=np.uint16)
        keep[kk] |= 0x00C0
        keep[kk & ~unmap] |= 0x000B
        keep[sec & supp] &= 0xF7FF
        assert ((y & 0xFFEF) == (w & keep & 0xFFEF))[sel].all()

        f1 = _rn16((w & 0x0200).view(np.float16).astype(np.float64) * _f(0x7800)).astype(np.float64)
        yf = _rn16(y.view(np.float16).astype(np.float64) * f1).view(np.uint16)
        assert (yf == np.where((w & 0x0200) != 0, y, 0)).all()
