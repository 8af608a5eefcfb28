"""Host-side mirror of the reference's TestData: identical synthetic data, identical time-level rotation."""
import numpy as np

from oracle import harness
from tinman_sandbox_b200.testdata import TestData


def test_init_data_bit_identical_to_reference_init(port):
    for E, L in ((6, 72), (3, 128), (2, 8)):
        want = port.init(E, L)
        got = TestData(E, L).init_data()
        for n in harness.FIELD_NAMES:
            assert np.array_equal(got.arrays[n], want.arrays[n]), (n, E, L)
        assert np.array_equal(got.ctl, want.ctl) and got.dt2 == want.dt2 and got.ps0 == want.ps0
        assert np.array_equal(got.consts, want.consts)
        assert np.array_equal(got.dvv, want.dvv) and np.array_equal(got.hyai, want.hyai)


def test_init_data_elem_offset_is_a_slice_of_the_global_init(port):
    want = port.init(10)
    got = TestData(4).init_data(elem_offset=3)
    for n in harness.FIELD_NAMES:
        assert np.array_equal(got.arrays[n], want.arrays[n][3:7]), n


def test_update_time_levels():
    td = TestData(1, 8).init_data()
    td.update_time_levels()      # reference: np1 <- nm1, nm1 <- n0, n0 <- old np1
    assert [int(x) for x in td.ctl[2:5]] == [1, 2, 0]
    td.update_time_levels()
    td.update_time_levels()
    assert [int(x) for x in td.ctl[2:5]] == [0, 1, 2]
