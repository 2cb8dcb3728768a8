"""SAM text -> hits on the device (kombgpu_sam_parse) through the C ABI, against the Python
restatement of the reference tokeniser (oracle.tokenise_sam, src/graph.cpp:197-239 at -t 1): the cases of
tests/test_host_tokenizer.py plus the Q2 / Q3 edge cases, and the graph built from the device-resident hits
against the golden komb2 outputs."""
import numpy as np
import pytest

from conftest import komb2_case_names, load_komb2_case
from komb_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import komb_b200
    c = komb_b200.Context(0)
    yield c
    c.close()


def check_against_restatement(oracle, ctx, sams):
    exp = [h for s in sams for h in oracle.tokenise_sam(s, 1)]      # -t 1 semantics: no line loss
    with ctx.sam_parse(sams) as hits:
        c = hits.counts()
        rk, ut = hits.download()
        names = hits.names()
    assert c["n_hits"] == len(exp) == rk.shape[0] == ut.shape[0]
    assert [names[v] for v in ut] == [r for _, r in exp]             # unitig of every hit, by Name, in file order
    # same key string <=> same read id, hit by hit; ids dense, in order of first appearance
    key_of, seen = {}, []
    for kid, (kstr, _) in zip(rk.tolist(), exp):
        if kid not in key_of:
            seen.append(kid)
        assert key_of.setdefault(kid, kstr) == kstr
    assert len(set(key_of.values())) == len(key_of) == c["n_reads"]
    assert seen == list(range(len(seen)))
    assert c["n_unitigs"] == len(names) == len(set(names))
    return rk, ut, names


@pytest.mark.parametrize("name", komb2_case_names())
def test_sam_parse_matches_reference_semantics(ctx, oracle_mod, name):
    sam1, sam2, _ = load_komb2_case(name)
    check_against_restatement(oracle_mod, ctx, [sam1, sam2])


def test_vid_order_is_deterministic_sq_then_first_seen(ctx):
    sam1 = (b"@HD\tVN:1.6\n@SQ\tSN:uB\tLN:5\n@SQ\tSN:uA\tLN:5\n@SQ\tSN:unused\tLN:5\n"
            b"r1/1\t0\tuA\t1\t60\t4M\t*\t0\t0\tACGT\tIIII\n"
            b"r1/1\t256\tnotInHeader\t1\t60\t4M\t*\t0\t0\tACGT\tIIII\n"
            b"r2/1\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII\n")
    sam2 = b"r1/2\t0\tuB\t1\t60\t4M\t*\t0\t0\tACGT\tIIII\nr9/2\t0\tuA\t1\t60\t4M\t*\t0\t0\tACGT\tIIII\n"
    with ctx.sam_parse([sam1, sam2]) as hits:
        rk, ut = hits.download()
        assert hits.names() == [b"uB", b"uA", b"notInHeader"]      # @SQ order, then first appearance; no hit -> no vertex
        assert ut.tolist() == [1, 2, 0, 1]
        assert rk[0] == rk[1] == rk[2] != rk[3]                      # r1/1, r1/1, r1/2 share key "1/"
        assert hits.counts()["n_lines"] == 9


def test_tokenizer_quirks(ctx, oracle_mod):
    # Q2: first QNAME char dropped, key runs through '/', reads differing only in char 0 merge
    sam1 = b"Xread7/1\t0\tu1\t1\nYread7/1\t0\tu2\t1\nplain\t0\tu3\t1\n"
    sam2 = b"Zlain\t0\tu4\t1\n\t\tab\t\t0\t\tu5\t1\n"                # Q3: runs of tabs collapse
    rk, ut, names = check_against_restatement(oracle_mod, ctx, [sam1, sam2])
    assert rk[0] == rk[1] and rk[2] == rk[3] and names[ut[4]] == b"u5"
    # one-character QNAME -> empty key; '/' first -> empty key too: both reads share it
    rk, ut, _ = check_against_restatement(oracle_mod, ctx, [b"a\t0\tu1\t1\n/x\t0\tu2\t1\nq/\t0\tu3\n", b""])
    assert rk[0] == rk[1] != rk[2]
    # unterminated last line of EACH file is not processed (reference -t 1 behaviour); files do not run together
    with ctx.sam_parse([b"a/1\t0\tu1\t1\nb/1\t0\tu2\t1", b"c/1\t0\tu3\t1\nd/1\t0\tu4"]) as hits:
        assert hits.counts()["n_hits"] == 2 and hits.names() == [b"u1", b"u3"]
    # exactly three fields, '\r' kept inside the last field like the reference
    with ctx.sam_parse([b"a/1\t0\tu1\r\n"]) as hits:
        assert hits.names() == [b"u1\r"]
    # empty inputs
    with ctx.sam_parse([b"", b""]) as hits:
        assert hits.counts() == {"n_hits": 0, "n_reads": 0, "n_unitigs": 0, "n_lines": 0}
        with hits.build_graph() as g:
            assert g.counts() == (0, 0)
    with ctx.sam_parse([b"@HD\tVN:1.6\n@SQ\tSN:u\tLN:9\n"]) as hits:
        assert hits.counts()["n_hits"] == 0 and hits.names() == []


def test_malformed_input_is_rejected(ctx):
    import komb_b200
    with pytest.raises(komb_b200.KombGpuError, match="fewer than 3"):
        ctx.sam_parse([b"onlyonefield\n", b""])
    with pytest.raises(komb_b200.KombGpuError, match=r"input 1, line 2\): empty line"):
        ctx.sam_parse([b"a\t0\tu1\n", b"a\t0\tu1\n\nb\t0\tu2\n"])
    # the context stays usable
    with ctx.sam_parse([b"a\t0\tu1\n"]) as hits:
        assert hits.counts()["n_hits"] == 1


def test_large_random_sam(ctx, oracle_mod):
    s1, s2, _, _ = synth.tiny_sam_pair(seed=21, n_unitigs=5000, n_reads=20000, max_hits=3)
    _, _, names = check_against_restatement(oracle_mod, ctx, [s1, s2])
    assert names == sorted(set(names), key=lambda x: int(x))        # @SQ order = numeric order in the generator


@pytest.mark.parametrize("name", [n for n in komb2_case_names() if not n.endswith("_t4")])
def test_graph_from_device_hits_matches_golden(ctx, name):
    """SAM bytes -> device hits -> graph -> coreness, nothing tokenised on the host: the reference's outputs by Name."""
    sam1, sam2, exp = load_komb2_case(name)
    with ctx.sam_parse([sam1, sam2]) as hits:
        names = [n.decode() for n in hits.names()]
        with hits.build_graph() as g:
            u, v = g.edges()
            deg, core = g.degree(), g.coreness()
    got_edges = {tuple(sorted((names[a], names[b]))) for a, b in zip(u.tolist(), v.tolist())}
    assert got_edges == {tuple(sorted(e)) for e in exp["edges"]}
    assert {names[i]: (int(core[i]), int(deg[i])) for i in range(len(names))} == exp["kcore"]


def test_many_distinct_strings_and_long_names(ctx, oracle_mod):
    """Interning at a size where table probing, first-appearance ranks and byte comparison all matter."""
    rng = np.random.default_rng(5)
    n_reads, n_unitigs = 60000, 30000
    unitig_names = [b"unitig_with_a_long_prefix_%d_%s" % (i, b"x" * int(rng.integers(0, 40))) for i in range(n_unitigs)]
    lines1, lines2 = [b"@SQ\tSN:%s\tLN:100" % nm for nm in unitig_names[::7]], []
    for r in rng.permutation(n_reads):
        for lines, mate in ((lines1, 1), (lines2, 2)):
            for _ in range(int(rng.integers(0, 3))):
                lines.append(b"Rread%d/%d\t0\t%s\t1\t60" % (r, mate, unitig_names[int(rng.integers(0, n_unitigs))]))
    s1, s2 = b"\n".join(lines1) + b"\n", b"\n".join(lines2) + b"\n"
    check_against_restatement(oracle_mod, ctx, [s1, s2])


# ---- output files formatted on the device (row N2) --------------------------------------------------------------

def test_percent_f_matches_glibc(ctx):
    """"%f" on the device = six decimals of the exact binary value, ties to even (Python's % formatting is the same rule)."""
    rng = np.random.default_rng(3)
    xs = np.concatenate([
        np.array([0.0, 1.0 / 128, 3.0 / 128, 0.5, 0.9999995, 0.99999949999, 1e-7, 4.9e-324, 2.5e-7, 5e-7, 1.5e-6, 22.180709777918249,
                  123456.7890125, 8796093022207.5, 1e-300, 0.1, 0.2 + 0.1]),
        rng.random(20000) * 25.0, rng.random(2000) * 1e-5, np.ldexp(rng.integers(1, 1 << 20, 2000).astype(np.float64), -rng.integers(1, 30, 2000)),
    ])
    got = ctx.format_corea(xs).decode().split("\n")
    assert got[-1] == "" and len(got) == xs.shape[0] + 1
    for i, x in enumerate(xs.tolist()):
        assert got[i] == "%d\t%f" % (i, x), (i, x)
    import komb_b200
    for bad in (np.inf, np.nan, 2.0 ** 43):
        with pytest.raises(komb_b200.KombGpuError):
            ctx.format_corea(np.array([1.0, bad]))
    assert ctx.format_corea(np.zeros(0)) == b""


@pytest.mark.parametrize("name", ["mid_s2", "quickstart_s1", "wide_s4"])
def test_device_formatted_files(ctx, name):
    """The three files' bytes from the device = what the host would print row by row from the downloaded arrays."""
    sam1, sam2, _ = load_komb2_case(name)
    with ctx.sam_parse([sam1, sam2]) as hits:
        names = hits.names()
        with hits.build_graph() as g:
            u, v = g.edges()
            deg, core, score = g.degree(), g.coreness(), g.corea()
            assert g.format(g.FILE_EDGELIST) == b"".join(b"%d\t%d\n" % (a, b) for a, b in zip(u.tolist(), v.tolist()))
            exp_k = b"#VID\tName\tCoreness\tDegree\n" + b"".join(
                b"%d\t%s\t%d\t%d\n" % (i, names[i], int(core[i]), int(deg[i])) for i in range(len(names)))
            assert g.format(g.FILE_KCORE, hits) == exp_k
            assert g.format(g.FILE_COREA) == b"".join(b"%d\t%f\n" % (i, float(s)) for i, s in enumerate(score))
    # kcore.tsv needs the names (the hits the graph was built from); the other two files do not
    import komb_b200
    with ctx.build_graph(np.array([0, 0, 1, 1], np.uint32), np.array([0, 1, 1, 2], np.uint32), 3) as g2:
        g2.coreness()
        with pytest.raises(komb_b200.KombGpuError, match="names"):
            g2.format(g2.FILE_KCORE)
        with pytest.raises(komb_b200.KombGpuError, match="corea"):
            g2.format(g2.FILE_COREA)
        assert g2.format(g2.FILE_EDGELIST) == b"0\t1\n1\t2\n"
