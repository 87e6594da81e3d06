#!/usr/bin/env python3
"""Fixture generator (runs only where /root/reference exists): the three hand-calculated log-likelihoods of
src/test/dr/evomodel/substmodel/MsatFullLikelihoodTest.java:60-190 -- microsatellite data (3 and 4 states, one pattern),
default AsymmetricQuadraticModel (= stepwise mutation model, dr/oldevomodel/substmodel/AsymmetricQuadraticModel.java:146-172,
normalised by ComplexSubstitutionModel.setupMatrix :325-340), stationary (uniform) root frequencies -- which the reference
asserts to 1e-10.  Writes tests/golden/msat.json."""
import json, re, sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def main():
    src = open(f"{REF}/src/test/dr/evomodel/substmodel/MsatFullLikelihoodTest.java").read()
    sizes = [int(m) for m in re.findall(r"new Microsatellite\(1,\s*(\d+)\)", src)]
    patterns = [[int(v) for v in m.split(",")] for m in re.findall(r"addPattern\(new int\[\]\{([\d,\s]+)\}\)", src)]
    newicks = re.findall(r'new NewickImporter\(\s*"([^"]+)"\)', src)
    values = [float(v) for v in re.findall(r"double logL\d = (-[\d.]+);", src)]
    assert len(sizes) == len(patterns) == len(newicks) == len(values) == 3, (sizes, patterns, newicks, values)
    out = {"source": "src/test/dr/evomodel/substmodel/MsatFullLikelihoodTest.java:60-190 (assertEquals tolerance 1e-10)",
           "cases": [{"stateCount": s, "pattern": p, "newick": n, "logL": v}
                     for s, p, n, v in zip(sizes, patterns, newicks, values)]}
    json.dump(out, open(__file__.rsplit("/", 1)[0] + "/msat.json", "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
