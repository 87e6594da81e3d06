#!/usr/bin/env python3
"""Fixture generator (runs only where /root/reference exists).

Parses the alignments of the reference's own benchmark inputs
  examples/Benchmarks/benchmark1.xml  (1441 taxa x 987 nt, HKY kappa=2, no gamma; header comment "npatterns=593")
  examples/Benchmarks/benchmark2.xml  (62 taxa x 10869 nt, GTR + Gamma4 alpha=0.5; "npatterns=5565")
into unique site patterns (state codes of src/dr/evolution/datatype/Nucleotides.java:73-91) and writes
tests/golden/benchmark{1,2}_patterns.npz (uint8 states [taxa][patterns] + float64 weights).  The pattern counts in
the XML comments are asserted, which pins the pattern compression itself to the reference.
"""
import re, sys
import xml.etree.ElementTree as ET
import numpy as np

sys.path.insert(0, __file__.rsplit("/tests/", 1)[0])
import beast_mcmc_b200  # noqa
from harness import evomodel as em

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = __file__.rsplit("/", 1)[0]

for name, npat in (("benchmark1", 593), ("benchmark2", 5565)):
    path = f"{REF}/examples/Benchmarks/{name}.xml"
    root = ET.parse(path).getroot()
    aln = root.find("alignment")
    seqs = ["".join(s.itertext()).replace("\n", "").replace("\t", "").replace(" ", "") for s in aln.findall("sequence")]
    taxa = [s.find("taxon").get("idref") for s in aln.findall("sequence")]
    assert len(set(map(len, seqs))) == 1, set(map(len, seqs))
    states = em.encode_nucleotides(seqs)
    pats = em.Patterns.fromAlignment(states)
    comment = re.search(r"npatterns=(\d+)", open(path).read())
    print(name, "taxa", len(seqs), "sites", len(seqs[0]), "patterns", pats.patternCount, "xml says", comment and comment.group(1))
    assert pats.patternCount == npat, (pats.patternCount, npat)
    np.savez_compressed(f"{HERE}/{name}_patterns.npz", states=pats.states.astype(np.uint8), weights=pats.weights,
                        taxa=np.array(taxa))
