"""CPU: the reference arm of bench.py (`--impl reference`) honours the driver's contract -- one JSON line with the
metric / config of the GPU arm, `impl: reference`, a cpu_baseline that says what was timed, an e2e object without copies --
and never maps the product library (the arm is the oracle port on the host cores, nothing else)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_and_purity():
    code = (
        "import sys, json; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--genomes', '2'];"
        "import bench; rc = bench.main();"
        "libs = sorted({l.split()[-1] for l in open('/proc/self/maps') if 'libgrm' in l});"
        "print('LIBS ' + json.dumps(libs)); sys.exit(rc)"
    )
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    line = json.loads(next(l for l in lines if l.startswith("{")))
    libs = json.loads(next(l for l in lines if l.startswith("LIBS "))[5:])
    assert libs and all(os.path.basename(l) == "libgrmoracle.so" for l in libs), libs      # oracle only, no libgrmkm.so
    assert line["impl"] == "reference" and line["unit"] == "Gbases/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("Gbases/s k-mer matrix build")
    assert line["value"] > 0 and line["steps"] == 1 and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "genomes 0..1" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["config"] == "c2" and line["config"]["genomes"] == 2 and line["config"]["k"] == 31


def test_rank_other_than_zero_does_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""
