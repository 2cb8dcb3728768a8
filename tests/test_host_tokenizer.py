"""Host logic without a GPU: the komb2 SAM tokeniser/interner (host/sam_tokenizer.hpp,
via bin/komb2_tokenize) against the Python restatement of the reference tokeniser, and
the C ABI library's symbol table."""
import ctypes
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import komb2_case_names, load_komb2_case
from komb_b200 import synth

ROOT = Path(__file__).resolve().parents[1]
TOKENIZE = ROOT / "bin" / "komb2_tokenize"


def run_tokenize(tmp_path, sam1, sam2, threads):
    (tmp_path / "r1.sam").write_bytes(sam1)
    (tmp_path / "r2.sam").write_bytes(sam2)
    cp = subprocess.run([str(TOKENIZE), str(threads), str(tmp_path / "r1.sam"), str(tmp_path / "r2.sam")],
                        capture_output=True, text=True)
    if cp.returncode != 0:
        raise RuntimeError(cp.stderr)
    lines = cp.stdout.split("\n")
    n = int(lines[0].split()[1])
    names = lines[1:1 + n]
    h = int(lines[1 + n].split()[1])
    hits = [tuple(map(int, l.split("\t"))) for l in lines[2 + n:2 + n + h]]
    return names, hits


def canonical(hits_named):
    """{frozenset of unitig names per read key} as a sorted multiset: independent of id numbering."""
    by = {}
    for k, nm in hits_named:
        by.setdefault(k, set()).add(nm)
    return sorted(tuple(sorted(s)) for s in by.values())


@pytest.fixture(scope="module", autouse=True)
def built():
    if not TOKENIZE.exists():
        subprocess.run(["make", "-C", str(ROOT), "bin/komb2_tokenize"], check=True, capture_output=True)


@pytest.mark.parametrize("name", komb2_case_names())
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_tokenizer_matches_reference_semantics(tmp_path, oracle_mod, name, threads):
    sam1, sam2, _ = load_komb2_case(name)
    names, hits = run_tokenize(tmp_path, sam1, sam2, threads)
    exp = oracle_mod.tokenise_sam(sam1, 1) + oracle_mod.tokenise_sam(sam2, 1)   # -t 1 semantics, any thread count
    assert len(hits) == len(exp)
    got_named = [(k, names[v]) for k, v in hits]
    exp_named = [(k, r.decode()) for k, r in exp]
    assert canonical(got_named) == canonical(exp_named)
    # same key string <=> same key id, hit by hit
    key_of = {}
    for (kid, _), (kstr, _) in zip(hits, exp):
        assert key_of.setdefault(kid, kstr) == kstr
    assert len(set(key_of.values())) == len(key_of)
    assert [names[v] for _, v in hits] == [r.decode() for _, r in exp]


def test_vid_order_is_deterministic_sq_then_first_seen(tmp_path):
    sam1 = (b"@HD\tVN:1.6\n@SQ\tSN:uB\tLN:5\n@SQ\tSN:uA\tLN:5\n@SQ\tSN:unused\tLN:5\n"
            b"r1/1\t0\tuA\t1\t60\t4M\t*\t0\t0\tACGT\tIIII\n"
            b"r1/1\t256\tnotInHeader\t1\t60\t4M\t*\t0\t0\tACGT\tIIII\n"
            b"r2/1\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII\n")
    sam2 = b"r1/2\t0\tuB\t1\t60\t4M\t*\t0\t0\tACGT\tIIII\nr9/2\t0\tuA\t1\t60\t4M\t*\t0\t0\tACGT\tIIII\n"
    for t in (1, 4):
        names, hits = run_tokenize(tmp_path, sam1, sam2, t)
        assert names == ["uB", "uA", "notInHeader"]            # @SQ order, then first appearance; no hit -> no vertex
        assert [v for _, v in hits] == [1, 2, 0, 1]
        assert hits[0][0] == hits[1][0] == hits[2][0] != hits[3][0]   # r1/1, r1/1, r1/2 share key "1/"


def test_tokenizer_quirks(tmp_path):
    # Q2: first QNAME char dropped, key runs through '/', reads differing only in char 0 merge
    sam1 = b"Xread7/1\t0\tu1\t1\nYread7/1\t0\tu2\t1\nplain\t0\tu3\t1\n"
    sam2 = b"Zlain\t0\tu4\t1\n\t\tab\t\t0\t\tu5\t1\n"            # Q3: runs of tabs collapse
    names, hits = run_tokenize(tmp_path, sam1, sam2, 2)
    assert hits[0][0] == hits[1][0]            # "read7/" twice
    assert hits[2][0] == hits[3][0]            # "lain" from plain / Zlain
    assert names[hits[4][1]] == "u5"
    # unterminated last line is not processed (reference -t 1 behaviour)
    names, hits = run_tokenize(tmp_path, b"a/1\t0\tu1\t1\nb/1\t0\tu2\t1", b"", 1)
    assert len(hits) == 1
    # malformed input is rejected, not guessed
    with pytest.raises(RuntimeError):
        run_tokenize(tmp_path, b"onlyonefield\n", b"", 1)
    with pytest.raises(RuntimeError):
        run_tokenize(tmp_path, b"a\t0\tu1\n\nb\t0\tu2\n", b"", 1)


def test_large_random_sam_many_threads(tmp_path, oracle_mod):
    s1, s2, _, _ = synth.tiny_sam_pair(seed=21, n_unitigs=5000, n_reads=20000, max_hits=3)
    names, hits = run_tokenize(tmp_path, s1, s2, 8)
    exp = oracle_mod.tokenise_sam(s1, 1) + oracle_mod.tokenise_sam(s2, 1)
    assert [names[v] for _, v in hits] == [r.decode() for _, r in exp]
    assert names == sorted(set(names), key=lambda x: int(x))     # @SQ order = numeric order in the generator


def test_c_abi_exports_every_declared_symbol():
    """include/*.h <-> libkombgpu.so <-> the ctypes table (no compute call without a GPU)."""
    from komb_b200 import _lib
    header = "".join(p.read_text() for p in sorted((ROOT / "include").glob("*.h")))
    declared = set(re.findall(r"\b(kombgpu_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.kombgpu_abi_version() == 1
    assert ctypes.sizeof(_lib.Stats) == 80   # sizeof(kombgpu_stats) on LP64


def test_komb2_cli_and_no_cpu_fallback(tmp_path):
    """The drop-in's command line (reference src/komb2.cpp:28-76, TCLAP semantics) needs no GPU; the analysis does, and
    without one komb2 fails loudly (exit 1, message on stderr) -- for the device tokeniser and for the host one alike:
    there is no CPU fallback."""
    import os
    komb2 = ROOT / "bin" / "komb2"
    if not komb2.exists():
        subprocess.run(["make", "-C", str(ROOT), "bin/komb2"], check=True, capture_output=True)
    cp = subprocess.run([str(komb2), "--version"], capture_output=True, text=True)
    assert cp.returncode == 0 and "version: 2.0" in cp.stdout
    cp = subprocess.run([str(komb2), "-i", "x.sam"], capture_output=True, text=True)
    assert cp.returncode == 1 and "Required argument" in cp.stderr
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is visible: the no-fallback half of this test is for CPU-only hosts")
    except ImportError:
        pass
    (tmp_path / "r1.sam").write_bytes(b"r1/1\t0\tu1\t1\nr1/1\t0\tu2\t1\n")
    (tmp_path / "r2.sam").write_bytes(b"r1/2\t0\tu3\t1\n")
    (tmp_path / "u.fa").write_text(">u1\nACGT\n")
    for mode in ("gpu", "host"):
        cp = subprocess.run([str(komb2), "-t", "2", "-i", str(tmp_path / "r1.sam"), "-j", str(tmp_path / "r2.sam"), "-u", str(tmp_path / "u.fa"),
                             "-o", str(tmp_path)], capture_output=True, text=True, env={**os.environ, "KOMB_TOKENIZE": mode})
        assert cp.returncode == 1 and "cannot use CUDA device" in cp.stderr, (mode, cp.stderr)
        assert not (tmp_path / "kcore.tsv").exists()


def test_python_api_fails_loudly_without_a_gpu():
    """komb_b200.Context raises KOMBGPU_ENODEV on a host without a usable B200: nothing in the package computes on the CPU."""
    import komb_b200
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is visible")
    except ImportError:
        pass
    with pytest.raises(komb_b200.KombGpuError) as e:
        komb_b200.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
