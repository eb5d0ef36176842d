#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference tree (run in the build container only;
/root/reference does not exist on the GPU box, the committed fixtures travel instead).

1. copies the reference's own fixtures for the sketch/dist path (data files, not sources);
   large text inputs are gzipped with mtime 0 so the bytes are reproducible;
2. generates known-answer vectors by calling the reference's own hash.cpp / MinHashHeap.cpp
   (oracle/_ref/libmashref.so) on seeded inputs -> ref_vectors.json;
3. runs the reference's own lyn2vec (README.md:34-52 recipe) on DNA1.fasta with --type_factorization ICFL,
   CFL_ICFL-30, CFL_COMB, ICFL_COMB and CFL_ICFL_COMB-10 -> DNA1-ICFL.txt.gz, DNA1-CFL_ICFL-30.txt.gz, DNA1-CFL_COMB.txt.gz ... (the CFL run reproduces the shipped DNA1-CFL.txt
   byte for byte, which is checked here).
"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle"))

COPY = {
    "mash/test_sequence.msh": "test_sequence.msh",
    "mash/new_data/test_sequence.fasta": "test_sequence.fasta",
    "mash/new_data/reads/reads.msh": "reads.msh",
    "mash/test/ref/reads.json": "reads.json",
    "mash/test/ref/genomes.dist": "genomes.dist",
    "mash/data/genome1.fna.msh": "genome1.fna.msh",
    "mash/data/genome2.fna.msh": "genome2.fna.msh",
    "mash/data/genome3.fna.msh": "genome3.fna.msh",
    "mash/test/paste_example/read1_2.msh": "read1_2.msh",
    "mash/example1.msh": "example1.msh",
    "training/Umberto/CFL/DNA1-sketch.msh": "DNA1-sketch.msh",
    "training/Umberto/CFL/DNA2-sketch.msh": "DNA2-sketch.msh",
    "training/Umberto/CFL/DNA3-sketch.msh": "DNA3-sketch.msh",
    "training/Umberto/CFL/DNA1.fasta": "DNA1.fasta",
}
GZIP = {
    "mash/new_data/reads/reads1.fastq": "reads1.fastq.gz",
    "mash/new_data/reads/reads2.fastq": "reads2.fastq.gz",
    "training/Umberto/CFL/DNA1-CFL.txt": "DNA1-CFL.txt.gz",
    "training/Umberto/CFL/DNA2-CFL.txt": "DNA2-CFL.txt.gz",
    "training/Umberto/CFL/DNA3-CFL.txt": "DNA3-CFL.txt.gz",
    "training/Umberto/CFL/DNA1-sketch.json": "DNA1-sketch.json.gz",
}


def main():
    for src, dst in COPY.items():
        shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
        os.chmod(os.path.join(HERE, dst), 0o644)
    for src, dst in GZIP.items():
        with open(os.path.join(REF, src), "rb") as f, open(os.path.join(HERE, dst), "wb") as raw:
            with gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0, compresslevel=9) as g:
                g.write(f.read())

    # lyn2vec, run as the README runs it (it writes fingerprint_<type>.txt next to the FASTA file)
    with tempfile.TemporaryDirectory() as d:
        shutil.copyfile(os.path.join(REF, "training/Umberto/CFL/DNA1.fasta"), os.path.join(d, "DNA1.fasta"))
        for fact in ("CFL", "ICFL", "CFL_ICFL-30", "CFL_COMB", "ICFL_COMB", "CFL_ICFL_COMB-10"):
            subprocess.run([sys.executable, "lyn2vec.py", "--type", "basic", "--path", d + "/", "--fasta", "DNA1.fasta", "--type_factorization", fact,
                            "--rev_comb", "true", "-n", "4"], cwd=os.path.join(REF, "lyn2vec"), check=True, stdout=subprocess.DEVNULL)
            data = open(os.path.join(d, "fingerprint_%s.txt" % fact), "rb").read()
            if fact == "CFL":
                assert data == open(os.path.join(REF, "training/Umberto/CFL/DNA1-CFL.txt"), "rb").read(), "lyn2vec no longer reproduces DNA1-CFL.txt"
                continue
            with open(os.path.join(HERE, "DNA1-%s.txt.gz" % fact), "wb") as raw:
                with gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0, compresslevel=9) as g:
                    g.write(data)

    from oracle_py import RefLib, build
    build(ref=True)
    ref = RefLib()
    rng = np.random.default_rng(20261018)
    vec = {"hash": [], "fp": [], "heap": [], "sketch": []}
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    for k in range(1, 33):
        for seed in (42, 0, 4294967295):
            s = lut[rng.integers(0, 4, size=k)].tobytes()
            vec["hash"].append({"kmer": s.decode(), "seed": seed, "h64": str(ref.get_hash(s, seed, True)), "h32": ref.get_hash(s, seed, False)})
    for n in list(range(0, 12)) + [17, 33]:
        toks = [int(x) for x in rng.integers(0, 1000, size=n)]
        vec["fp"].append({"tokens": toks, "seed": 42, "h32": ref.fp_hash(toks, 42, False), "h64": str(ref.fp_hash(toks, 42, True))})
    for t in range(40):
        s = int(rng.integers(1, 9)); m = int(rng.integers(1, 4)); use64 = bool(t & 1)
        stream = rng.integers(0, 25, size=int(rng.integers(0, 120)))
        h, c, est = ref.heap_stream(stream, s, m, use64)
        vec["heap"].append({"s": s, "m": m, "use64": use64, "stream": [int(x) for x in stream], "hashes": [int(x) for x in h],
                            "counts": [int(x) for x in c], "set_size": est})
    for t, (k, s, m) in enumerate([(21, 200, 1), (16, 100, 1), (32, 150, 1), (21, 50, 2), (7, 1000, 1), (21, 64, 3)]):
        recs = []
        for r in range(3):
            x = bytearray(lut[rng.integers(0, 4, size=int(rng.integers(10, 4000)))].tobytes())
            for i in rng.integers(0, len(x), size=5):
                x[i] = ord("N")
            for i in rng.integers(0, len(x), size=40):
                x[i] |= 0x20
            recs.append(bytes(x))
        if m > 1:
            recs = recs + recs[:2]
        out = ref.sketch(recs, k=k, s=s, min_cov=m)
        vec["sketch"].append({"k": k, "s": s, "min_cov": m, "records": [r.decode() for r in recs], "hashes": [str(int(x)) for x in out["hashes"]],
                              "counts": [int(x) for x in out["counts"]], "set_size": out["set_size"]})
    with open(os.path.join(HERE, "ref_vectors.json"), "w") as f:
        json.dump(vec, f, indent=0)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
