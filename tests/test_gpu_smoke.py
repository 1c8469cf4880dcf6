import pytest


@pytest.mark.gpu
def test_smoke_bit_exact():
    import __graft_entry__ as ge
    ge.smoke()
