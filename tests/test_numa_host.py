"""CPU-only: the NUMA helper degrades gracefully (no GPU, no nvidia-smi, single node) and never widens the affinity mask."""
import os


def test_topology_and_bind_report_without_gpu():
    from pplp_b200 import numa
    before = os.sched_getaffinity(0)
    topo = numa.topology()
    assert set(topo["cpus_allowed"]) == set(before)
    assert all(isinstance(v, int) for v in topo["nodes"].values())
    rep = numa.bind_to_gpu_node(0)
    assert rep["gpu"] == 0 and "bound" in rep
    assert os.sched_getaffinity(0) <= before          # binding can only narrow the mask
    os.sched_setaffinity(0, before)


def test_cpulist_parser():
    from pplp_b200.numa import _parse_cpulist
    assert _parse_cpulist("0-3,8,10-11") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("") == set() and _parse_cpulist(None) == set()
