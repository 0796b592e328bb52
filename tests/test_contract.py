"""Repository contracts the driver and the judge rely on (CPU only): the product never touches the oracle,
bench.py's reference arm prints one well-formed JSON line, build() leaves the library in-tree."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_sources_never_reference_the_oracle():
    # oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "astr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"pyoracle|oracle/|astr_oracle|libastr_oracle", txt):
                    bad.append(os.path.join(base, f))
    assert not bad, bad
    hdr = open(os.path.join(ROOT, "include", "astr_gpu.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)       # declarations only, comments stripped
    assert "torch" not in code and "at::" not in code and "std::" not in code   # plain C ABI: pointers, ints, doubles


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--cpu-n", "32"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mpts/s" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1", "--cpu-n", "32"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_built_library_is_in_tree_and_git_ignored():
    import astr_b200
    so = astr_b200.build()
    assert os.path.commonpath([so, ROOT]) == ROOT and os.path.exists(so)
    ig = open(os.path.join(ROOT, ".gitignore")).read()
    assert "*.so" in ig and "oracle/_ref/" in ig
