"""Arithmetic known-answer tests the kernels' design rests on (SURVEY.md sections 7.3, 7.4, Appendix B)."""
import os

import numpy as np
import pytest


def test_bresenham_closed_form_all_endpoints_to_401(oracle):
    """cell k = (major0 + k*s, minor0 + s'*floor((k*n + m/2)/m)) for every |dx|,|dy| <= 401, all octants."""
    assert oracle.L.orc_bresenham_closed_form_check(401) == 0


def test_magic_division_exact_up_to_1024_cells(oracle):
    assert oracle.L.orc_magic_division_check() == 0


def test_clamp_add_monoid(oracle):
    assert oracle.L.orc_clamp_monoid_check(20000, 12345) == 0


def test_time_slice_map_representation(oracle):
    """(f(lo_min), f(lo_max), saturating sum) reproduces any slice's effect on any start value, and composes."""
    assert oracle.L.orc_slice_map_check(3000, 7) == 0


def test_saturating_updates_do_not_commute():
    """78 +6 -1 = 79 but 78 -1 +6 = 80 -- why the kernels never reorder a cell's updates."""
    c = lambda v: max(-80, min(80, v))
    assert c(c(78 + 6) - 1) == 79 and c(c(78 - 1) + 6) == 80


def test_sincosf_restatement_equals_glibc_exhaustive(oracle):
    """EVERY float (4.29 G bit patterns: |y| < 120, glibc's reduce_large branch above that, Inf and all NaNs):
    restated glibc-2.39 sincosf == libm sincosf, bit for bit."""
    threads = min(os.cpu_count() or 1, 64)
    bad, first = oracle.sincosf_sweep(0, 0x42F00000, 1, threads)
    assert bad == 0, f"|y| < 120: {bad} mismatches, first at bits {first:#x}"
    bad, first = oracle.sincosf_sweep(0x42F00000, 0x80000000, 1, threads)
    assert bad == 0, f"|y| >= 120 / Inf / NaN: {bad} mismatches, first at bits {first:#x}"


def test_glibc_sincosf_is_not_correctly_rounded(oracle):
    """the host result differs from (float)sin((double)x) -- a correctly rounded device sinf would not match."""
    rng = np.random.default_rng(1)
    a = rng.uniform(-7, 7, 2_000_000).astype(np.float32)
    s, c = oracle.libm_sincosf(a)
    cr = np.sin(a.astype(np.float64)).astype(np.float32)
    assert (s != cr).sum() > 0
