"""CPU test of the bench.py contract pieces that need no GPU: the reference arm (--impl reference) prints one JSON
line with the keys the driver reads, and the workload table names BASELINE.json's configs."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "candidate cuts scored+selected/sec" and line["unit"] == "subsets/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "blocks" in cb["sample"]
    assert line["e2e"] == dict(value=line["value"], unit="subsets/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert line["gpu_launches"] == 0 and line["config"]["workload"]
    # same config object as the GPU arm prints for this workload
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    assert line["config"] == bench.config_of(bench.WORKLOADS["small"], argparse.Namespace(strat=4, gpus=1))


def test_reference_arm_does_not_load_the_product():
    """The CPU arm may execute oracle/ only: neither the package nor libsdpcutsel.so may be loaded by it."""
    code = ("import sys, runpy\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'small', '--steps', '1', '--warmup', '1']\n"
            "runpy.run_path(%r, run_name='__main__')\n"
            "bad = [m for m in sys.modules if 'sdpcutsel' in m]\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert not bad and 'libsdpcutsel' not in maps, (bad, 'libsdpcutsel' in maps)\n" % os.path.join(ROOT, "bench.py"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    src = open(os.path.join(ROOT, "bench.py")).read()
    job = src[src.index("def cpu_block_job"):src.index("def sample_blocks")]
    assert "sdpcutsel_via_nn_b200" not in job and "pkg." not in job


def test_workloads_name_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "234" in bench.WORKLOADS["cfg4"]["name"] and bench.WORKLOADS["cfg4"]["n"] == 125 and bench.WORKLOADS["cfg4"]["rho"] == 5
    assert "C(125,5)" in base["configs"][3] and "C(125,4)" in base["configs"][2]
    assert bench.comb(125, 5) == 234531275 and bench.comb(125, 4) == 9691375
    assert bench.W_FLOPS[5] == 27264 + 1368 + 76
