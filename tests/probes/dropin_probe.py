"""Wall time of the komb2 drop-in against the reference binary on the same SAM pair (cfg2 read sample):
   tests/probes/dropin_probe.py [read_pairs=625000] [threads]"""
import os, re, subprocess, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from komb_b200 import synth
from oracle import oracle
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 625_000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
n = 1_000_000
m1, m2 = synth.metagenome_hits(n, pairs, seed=11)
with tempfile.TemporaryDirectory() as d:
    d = Path(d)
    (d / "r1.sam").write_bytes(synth.render_sam(m1, n, 1, with_header=False))
    (d / "r2.sam").write_bytes(synth.render_sam(m2, n, 2, with_header=False))
    synth.write_fasta(str(d / "u.fasta"), 4)
    print("hits", m1.n_hits + m2.n_hits, "SAM bytes", (d / "r1.sam").stat().st_size + (d / "r2.sam").stat().st_size, flush=True)
    for name, binary in (("komb2 (this repo)", ROOT / "bin" / "komb2"), ("komb2 (this repo), warm", ROOT / "bin" / "komb2"),
                         ("komb2_ref (reference)", oracle.REF_KOMB2)):
        if not Path(binary).exists():
            continue
        out = d / ("out_" + name.split()[0] + str(len(name)))
        out.mkdir()
        t0 = time.perf_counter()
        os.environ["KOMB_TIMING"] = "1"
        cp = subprocess.run([str(binary), "-t", str(threads), "-l", "100", "-o", str(out), "-i", str(d / "r1.sam"), "-j", str(d / "r2.sam"),
                             "-u", str(d / "u.fasta")], capture_output=True, text=True)
        wall = time.perf_counter() - t0
        stages = {m.group(1).strip(): float(m.group(2)) for m in re.finditer(r"Time elapsed (?:for|doing) ([^:]+): ([0-9.]+) s", cp.stdout)}
        ana = re.search(r"analysis \(sec\) = ([0-9.]+)", cp.stdout)
        print(f"{name}: rc={cp.returncode} wall {wall:.3f} s  analysis {ana.group(1) if ana else None}  stages {stages}", flush=True)
        if cp.returncode or "timing" in cp.stderr:
            print(cp.stderr[-1500:])
