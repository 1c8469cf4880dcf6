import pytest


@pytest.mark.gpu
def test_smoke_entry_point():
    import __graft_entry__ as ge
    ge.smoke()
