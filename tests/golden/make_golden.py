#!/usr/bin/env python3
"""Regenerates the golden fixtures in this directory from the reference checkout.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
The fixtures are the inputs + expected values of the reference's own known-answer tests for the
hot path; they are data, not code.
"""
import json
import re
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def hla_faux():
    # test_data/HLA-faux/database.json: the 2-allele DB behind test_reference_alleles
    # (src/hla/caller.rs:1709-1773) and test_score_bad_read (:1783-1809)
    db = json.loads((REF / "test_data/HLA-faux/database.json").read_text())
    alleles = {}
    for hla_id, d in db["hla_sequences"].items():
        alleles[hla_id] = dict(gene_name=d["gene_name"], star_allele=d["star_allele"],
                               dna_sequence=d["dna_sequence"], cdna_sequence=d["cdna_sequence"])
    expected = {
        # exact-copy read => best id is that allele with stats (cdna_len,0,0,dna_len,0,0)
        "test_reference_alleles": [
            dict(gene="HLA-A", hla_id="HLA:HLA00037", star="03:01:01:01", read_is_revcomp=False),
            dict(gene="HLA-B", hla_id="HLA:HLA00132", star="07:02:01:01", read_is_revcomp=True),
        ],
        # 4-bp read => no best match, every score is the worst (1.0); cDNA scoring disabled
        "test_score_bad_read": dict(gene="HLA-A", read="ACGT", worst_score=1.0),
    }
    (OUT / "hla_faux.json").write_text(json.dumps(dict(
        source="test_data/HLA-faux/database.json", database_metadata=db["database_metadata"],
        hla_sequences=alleles, expected=expected), indent=1) + "\n")


def weight_sequence():
    # src/cyp2d6/chaining.rs:1050-1080 (test_weight_sequence): three consensuses differing at one
    # base; query 1 copies consensus 0 => strictly best; query 2 has N there => all three tie
    src = (REF / "src/cyp2d6/chaining.rs").read_text().splitlines()[1049:1080]
    cons = [m.group(1) for line in src for m in [re.search(r'Consensus::new\(b"([ACGTN]+)"', line)] if m]
    queries = [m.group(1) for line in src for m in [re.search(r'let sequence = "([ACGTN]+)"', line)] if m]
    assert len(cons) == 3 and len(queries) == 2
    (OUT / "weight_sequence.json").write_text(json.dumps(dict(
        source="src/cyp2d6/chaining.rs:1050-1080", consensuses=cons, queries=queries,
        expected=["consensus 0 is the strict minimum", "all three scores are equal"]), indent=1) + "\n")


if __name__ == "__main__":
    hla_faux()
    weight_sequence()
    print("wrote", sorted(p.name for p in OUT.glob("*.json")))
