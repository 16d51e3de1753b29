"""Host-side NUMA placement helper (quantizers_b200/numa.py)."""

def test_numa_cpulist_parsing():
    from quantizers_b200 import numa

    assert numa.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert numa.parse_cpulist("") == set()
    assert numa.parse_cpulist("a-b") == set()
    # no GPU / unknown device: the context manager leaves the thread's affinity alone
    import os

    before = os.sched_getaffinity(0)
    with numa.near_device(0) as bound:
        assert bound in (False, True)
    assert os.sched_getaffinity(0) == before
