"""The drop-in executable: bin/komb2 run the way KOMB.py runs it (reference
KOMB.py:436-442) on the SAM pairs whose reference outputs are committed under
tests/golden/komb2 — the three output files must agree canonically (by unitig
Name, quirks Q4/Q7) with what komb2_ref -t 1 wrote."""
import subprocess
from pathlib import Path

import pytest

from conftest import komb2_case_names, load_komb2_case

ROOT = Path(__file__).resolve().parents[1]
KOMB2 = ROOT / "bin" / "komb2"

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", [n for n in komb2_case_names() if not n.endswith("_t4")])
@pytest.mark.parametrize("threads,tokenise,output", [(1, "gpu", "gpu"), (5, "gpu", "host"), (3, "host", "host")])
def test_komb2_outputs_match_reference(tmp_path, oracle_mod, name, threads, tokenise, output):
    """tokenise: SAM text parsed and interned on the device (kombgpu_sam_parse, the default) or by the host tokeniser;
    output: the three files formatted on the device (kombgpu_graph_format, the default) or by the host."""
    sam1, sam2, exp = load_komb2_case(name)
    got, stdout = oracle_mod.run_komb2(KOMB2, sam1, sam2, tmp_path, threads=threads,
                                       extra_env={"KOMB_TOKENIZE": tokenise, "KOMB_OUTPUT": output})
    assert got["edges"] == exp["edges"]
    assert got["kcore"] == exp["kcore"]
    assert set(got["score_text"]) == set(exp["score_text"])
    for nm, txt in exp["score_text"].items():
        # identical %f text unless the 7th decimal is a rounding tie (1-ulp log differences)
        assert got["score_text"][nm] == txt or abs(float(got["score_text"][nm]) - float(txt)) <= 1.0000001e-6
    for line in exp["stdout_info"]:
        if line.startswith("Max CoreA"):
            continue
        assert line in stdout
    for marker in ("Time elapsed for reading SAMs", "Time elapsed for generateGraph", "GraphInfo...",
                   "Time elapsed doing K-core decomposition", "Created Kcore", "Time elapsed for anomalyDetection",
                   "Time elapsed for KOMB"):
        assert marker in stdout


def test_komb2_t4_case_uses_t1_semantics(tmp_path, oracle_mod):
    """The reference at -t 4 drops lines (quirk Q1); the drop-in never does, whatever -t is."""
    sam1, sam2, exp_t4 = load_komb2_case("mid_s3_t4")
    got, _ = oracle_mod.run_komb2(KOMB2, sam1, sam2, tmp_path, threads=4)
    exp_t1 = oracle_mod.komb2_expected(sam1, sam2, threads=1)
    assert got["edges"] == exp_t1["edges"] and got["kcore"] == exp_t1["kcore"]
    assert exp_t4["edges"] < got["edges"]


def test_komb2_error_behaviour(tmp_path):
    cp = subprocess.run([str(KOMB2), "-i", "x.sam"], capture_output=True, text=True)
    assert cp.returncode == 1 and "Required argument" in cp.stderr
    cp = subprocess.run([str(KOMB2), "--version"], capture_output=True, text=True)
    assert cp.returncode == 0 and "version: 2.0" in cp.stdout
    cp = subprocess.run([str(KOMB2), "-i", str(tmp_path / "missing.sam"), "-j", "b", "-u", "c", "-o", str(tmp_path)],
                        capture_output=True, text=True)
    assert cp.returncode == 1 and "could not be opened" in cp.stderr
    # malformed SAM text is rejected with a message, by the device tokeniser and by the host one
    (tmp_path / "bad.sam").write_bytes(b"r1/1\t0\tu1\nonlyonefield\n")
    (tmp_path / "ok.sam").write_bytes(b"r1/2\t0\tu2\n")
    (tmp_path / "u.fa").write_text(">u1\nACGT\n")
    for mode in ("gpu", "host"):
        cp = subprocess.run([str(KOMB2), "-i", str(tmp_path / "bad.sam"), "-j", str(tmp_path / "ok.sam"), "-u", str(tmp_path / "u.fa"),
                             "-o", str(tmp_path)], capture_output=True, text=True, env={**__import__("os").environ, "KOMB_TOKENIZE": mode})
        assert cp.returncode == 1 and "malformed SAM" in cp.stderr and "fewer than 3" in cp.stderr


def test_komb2_exact64_key_mode(tmp_path, oracle_mod):
    sam1, sam2, _ = load_komb2_case("quickstart_s1")
    got, _ = oracle_mod.run_komb2(KOMB2, sam1, sam2, tmp_path, extra_env={"KOMB_COREA_KEY": "exact64"})
    exp = oracle_mod.komb2_expected(sam1, sam2, key_mode=oracle_mod.KEY_EXACT64)
    for nm, s in exp["score"].items():
        assert abs(got["score"][nm] - s) <= 5.0e-7 + 1e-12


def _device_list():
    """Distinct GPUs when the box has them, else every rank on device 0 (the library's emulation of several ranks on
    one device: the same partitioned path, host-side exchanges)."""
    import torch
    n = torch.cuda.device_count()
    return ",".join(str(d) for d in range(min(n, 4))) if n >= 2 else "0,0,0"


@pytest.mark.parametrize("name", ["mid_s2", "quickstart_s1", "wide_s4"])
def test_komb2_multi_gpu_outputs_match_reference(tmp_path, oracle_mod, name):
    """KOMB_GPU_DEVICES=...: the C++ host drives one rank per device (threads, peer-memory path, no torch, no NCCL);
    the three output files are the ones the reference writes."""
    sam1, sam2, exp = load_komb2_case(name)
    got, stdout = oracle_mod.run_komb2(KOMB2, sam1, sam2, tmp_path, threads=4, extra_env={"KOMB_GPU_DEVICES": _device_list()})
    assert got["edges"] == exp["edges"]
    assert got["kcore"] == exp["kcore"]
    for nm, txt in exp["score_text"].items():
        assert got["score_text"][nm] == txt or abs(float(got["score_text"][nm]) - float(txt)) <= 1.0000001e-6
    for line in exp["stdout_info"]:
        if not line.startswith("Max CoreA"):
            assert line in stdout
