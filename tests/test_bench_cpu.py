"""bench.py host logic that needs no GPU: the reference arm runs the oracle port only (it must not map the product .so),
prints a line whose `config` equals the GPU arm's for the same arguments, and `ms_per_step` is consistent with `value`."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_never_loads_the_product_library():
    code = (
        "import sys, runpy\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'nodes', '--steps', '2', '--warmup', '1', '--batch', '4']\n"
        "try:\n"
        "    runpy.run_path('bench.py', run_name='__main__')\n"
        "finally:\n"
        "    maps = open('/proc/self/maps').read()\n"
        "    sys.stderr.write('MAPPED_PRODUCT_SO=%d\\n' % maps.count('libcv_b200'))\n"
    )
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "MAPPED_PRODUCT_SO=0" in r.stderr
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["gpu_launches"] == 0 and line["higher_is_better"] is True
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    # value and ms_per_step describe the same steps
    per_step = 4
    assert abs(line["value"] - per_step / (line["ms_per_step"] / 1e3)) <= 1e-6 * line["value"]


def test_both_arms_print_the_same_config_object():
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")

    class A:
        batch, size, variant, total = 64, 1024, "tiny", 8192
    for wl in ("pipeline", "sam2", "nodes", "cfg5"):
        c1, c2 = bench.make_config(wl, A), bench.make_config(wl, A)
        assert c1 == c2 and "workload" in c1 and set(c1) >= {"images_per_gpu_per_step", "size", "l2_policy", "sharding"}
    assert "cv_ccl_label" not in bench.workload_name("nodes4096", A)  # the cfg-4 line times cv_nodes_analyze, and says so
