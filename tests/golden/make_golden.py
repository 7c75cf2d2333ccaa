#!/usr/bin/env python
"""Extracts the reference's own golden vectors for the hot path into JSON fixtures.

Run HERE (the only place /root/reference exists):  python tests/golden/make_golden.py
Reads (never copies code, only literal test data):
  * source/dentist/dazzler.d:965-1026   LAdump text fed to `dumpLA`
  * source/dentist/dazzler.d:1045-1166  expected FlatLocalAlignment[]  (ids 1-based, DENTIST flags)
  * source/dentist/dazzler.d:502-654    expected AlignmentChain[]      (chain packing)
  * source/dentist/common/alignments/base.d:886-941  real 21-tile daligner trace + 12 assertions
"""
import json
import os
import re

REF = "/root/reference/source/dentist"
HERE = os.path.dirname(os.path.abspath(__file__))


def lines(path, a, b):
    with open(path) as f:
        return f.read().split("\n")[a - 1:b]


def main():
    dz = os.path.join(REF, "dazzler.d")
    dump = [m.group(1) for ln in lines(dz, 965, 1026) for m in [re.search(r'"(.*)"', ln)] if m]
    txt = "\n".join(lines(dz, 1045, 1166))
    flat = []
    for m in re.finditer(
            r"FlatLocalAlignment\(\s*(\d+),\s*FlatLocus\((\d+), (\d+), (\d+), (\d+)\),\s*FlatLocus\((\d+), (\d+), (\d+), (\d+)\),"
            r"\s*AlignmentFlags\((.*?)\),\s*expectedTracePointDistance,\s*\[(.*?)\],\s*\)", txt, re.S):
        g = m.groups()
        flags = re.findall(r"AlignmentFlag\.(\w+)", g[9])
        tps = [[int(a), int(b)] for a, b in re.findall(r"TracePoint\((\d+), (\d+)\)", g[10])]
        flat.append(dict(id=int(g[0]), contigA=int(g[1]), abpos=int(g[3]), aepos=int(g[4]),
                         contigB=int(g[5]), bbpos=int(g[7]), bepos=int(g[8]), flags=flags, trace=tps))
    txt = "\n".join(lines(dz, 502, 654))
    chains = []
    for m in re.finditer(r"AlignmentChain\(\s*(\d+),\s*Contig\((\d+), 0\),\s*Contig\((\d+), 0\),\s*AlignmentFlags\((.*?)\),\s*\[(.*?)\],\s*expectedTracePointDistance",
                         txt, re.S):
        las = re.findall(r"LocalAlignment\(\s*Locus\((\d+), (\d+)\),\s*Locus\((\d+), (\d+)\),\s*(\d+),", m.group(5))
        chains.append(dict(id=int(m.group(1)), contigA=int(m.group(2)), contigB=int(m.group(3)),
                           flags=re.findall(r"AlignmentFlag\.(\w+)", m.group(4)),
                           las=[[int(x) for x in la] for la in las]))
    bd = os.path.join(REF, "common/alignments/base.d")
    txt = "\n".join(lines(bd, 886, 941))
    tps = [[int(a), int(b)] for a, b in re.findall(r"(?<!Translated)TracePoint\(\s*(\d+),\s*(\d+)\)", txt)]
    loci = re.findall(r"Locus\((\d+), (\d+)\)", txt)
    asserts = []
    for m in re.finditer(r"translateTracePoint\((\d+), RoundingMode\.(\w+)\) == TranslatedTracePoint\((\d+), ([0-9 +]+)\)", txt):
        asserts.append(dict(pos=int(m.group(1)), mode=m.group(2), a=int(m.group(3)), b=eval(m.group(4))))
    kat = dict(tspace=100, abpos=int(loci[0][0]), aepos=int(loci[0][1]), bbpos=int(loci[1][0]), bepos=int(loci[1][1]),
               diffs=292, trace=tps, asserts=asserts, throws=[578, 2585],
               equal_pairs=[[700, "ceil", 700, "floor"], [699, "ceil", 701, "floor"]])
    # consensus KAT dazzler.d:4257-4299: three 1050-bp reads, reads 1 and 2 carry one substitution each
    # (capital letter), `daligner -l15` + filterPileUpAlignments(allowance 100) + daccord => consensus == read 3
    fa = [m.group(1) for ln in lines(dz, 4262, 4266) for m in [re.search(r'"(>.*)"', ln)] if m]
    cons = dict(reads=["".join(r.split("\\n")[1:]) for r in fa], minlen=15, allowance=100, expected_read=2)
    with open(os.path.join(HERE, "consensus_kat.json"), "w") as f:
        json.dump(cons, f, indent=1)
    # getCroppingSlice KATs  commands/processPileUps/cropper.d:552-647
    cp = os.path.join(REF, "commands/processPileUps/cropper.d")
    txt = "\n".join(lines(cp, 552, 647))
    cases = []
    for blk in txt.split("enum alignment = SeededAlignment(")[1:]:
        contigs = re.findall(r"Contig\((\d+), (\d+)\)", blk)
        loci = re.findall(r"Locus\((\d+), (\d+)\)", blk)
        tps = [[int(a), int(b)] for a, b in re.findall(r"TracePoint\(\s*(\d+),\s*(\d+)\)", blk)]
        comp = "AlignmentFlags(complement)" in blk
        seed = re.search(r"AlignmentLocationSeed\.(\w+)", blk).group(1)
        asserts = [[int(p), int(b), int(e)] for p, b, e in
                   re.findall(r"ReferencePoint\(1, (\d+)\)\]\) ==\s*ReadInterval\(\d+, (\d+), (\d+)\)", blk)]
        cases.append(dict(read_len=int(contigs[1][1]), abpos=int(loci[0][0]), aepos=int(loci[0][1]), bbpos=int(loci[1][0]),
                          bepos=int(loci[1][1]), trace=tps, complement=comp, seed=seed, tspace=100, asserts=asserts))
    with open(os.path.join(HERE, "cropper_kat.json"), "w") as f:
        json.dump(cases, f, indent=1)
    # PileUpDb / InsertionDb test data  common/binio/_testdata/{pileupdb,insertiondb}.d (literal data only) with the
    # element counts the reference's size tests use (pileupdb.d:430-464, insertiondb.d:470-511)
    binio = {}
    for name in ("pileupdb", "insertiondb"):
        src = open(os.path.join(REF, "common/binio/_testdata/%s.d" % name)).read()
        counts = {m.group(1): int(m.group(2)) for m in re.finditer(r"enum (num\w+) = (\d+);", src)}
        body = src[src.index("return [", src.index("TestData()")) + len("return "):]
        body = body[:body.rindex("];") + 1]
        body = body.replace("CompressedSequence.from(", "Seq(").replace("AlignmentLocationSeed.", "")
        ns = dict(__builtins__={}, complement=1, front="front", back="back", begin="begin", end="end", pre="pre", post="post",
                  Contig=lambda a, b: [a, b], Locus=lambda a, b: [a, b], TracePoint=lambda a, b: [a, b], Seq=lambda x: x, id_t=lambda x: x,
                  AlignmentFlags=lambda *f: sum(f), ContigNode=lambda a, b: [a, b],
                  LocalAlignment=lambda a, b, d, t: dict(ab=a[0], ae=a[1], bb=b[0], be=b[1], diffs=d, trace=t),
                  AlignmentChain=lambda i, a, b, f, l, tpd=0: dict(id=i, contigA=a, contigB=b, flags=f, las=l, tpd=tpd),
                  SeededAlignment=lambda c, s: dict(c, seed=s), ReadAlignment=lambda *s: list(s),
                  InsertionInfo=lambda seq, cl, ov, rid: dict(sequence=seq, contig_length=cl, overlaps=ov, read_ids=rid),
                  Insertion=lambda s, e, info: dict(info, start=s, end=e))
        binio[name] = dict(counts=counts, data=eval(body, ns))
    with open(os.path.join(HERE, "binio_kat.json"), "w") as f:
        json.dump(binio, f, separators=(",", ":"))
    # BadAlignmentCoverageAssessor KATs  commands/maskRepetitiveRegions.d:302-340 (inputs), :395-411 (mask), :582-617 (changes)
    mp = os.path.join(REF, "commands/maskRepetitiveRegions.d")
    iv = lambda a, b: [[int(x) for x in m] for m in re.findall(r"ReferenceInterval\((\d+),\s*(\d+),\s*(\d+)\)", "\n".join(lines(mp, a, b)))]
    txt = "\n".join(lines(mp, 575, 650))
    maskcov = dict(alignments=iv(296, 331), contigs=iv(333, 341), bounds=[3, 5], mask=iv(402, 411),
                   changes=[[int(x) for x in m] for m in re.findall(r"CoverageChange\((\d+),\s*(\d+),\s*(\d+),\s*(\d+)\)", txt)])
    with open(os.path.join(HERE, "maskcov_kat.json"), "w") as f:
        json.dump(maskcov, f)
    # AlignmentChain.opCmp KAT  common/alignments/base.d:780-857: two lists that must come out strictly ascending
    bd2 = "\n".join(lines(os.path.join(REF, "common/alignments/base.d"), 780, 857))
    order_kat = []
    for blk in bd2.split("auto acs = [")[1:]:
        blk = blk[:blk.index("];")]
        la_default = [[0, 1], [0, 1]]
        chains_ = []
        for m in re.finditer(r"AlignmentChain\((\d+), Contig\((\d+), \d+\), Contig\((\d+), \d+\), Flags\(\), \[(.*?)\]\)", blk, re.S):
            las_ = re.findall(r"LocalAlignment\(Locus\((\d+), (\d+)\), Locus\((\d+), (\d+)\), \d+\)", m.group(4))
            las_ = [[int(x) for x in t] for t in las_] or [[0, 1, 0, 1]]
            chains_.append(dict(id=int(m.group(1)), contigA=int(m.group(2)), contigB=int(m.group(3)), las=las_))
        order_kat.append(chains_)
    with open(os.path.join(HERE, "chain_order_kat.json"), "w") as f:
        json.dump(order_kat, f)
    # ReadAlignment typing KAT  common/alignments/base.d:2324-2676: chains (seeds derived by SeededAlignment.from in the test)
    # and the expectation string per case: isInOrder isValid type isExtension isFront isBack isGap isParallel isAntiParallel
    src = lines(os.path.join(REF, "common/alignments/base.d"), 2329, 2604)
    body = "\n".join(ln for ln in src if not ln.strip().startswith("//"))
    body = body[body.index("["):].replace("SeededAlignment.from(", "SAfrom(")
    body = "{" + body[1:body.rindex("]")] + "}"
    class _Front(dict):
        front = property(lambda self: dict(self))
    ns = dict(__builtins__={}, complement=1, Contig=lambda a, b: [a, b], Locus=lambda a, b: [a, b], Flags=lambda *f: sum(f),
              LocalAlignment=lambda a, b, d: dict(ab=a[0], ae=a[1], bb=b[0], be=b[1], diffs=d),
              AlignmentChain=lambda i, a, b, f, l: _Front(id=i, contigA=a, contigB=b, flags=f, las=l),
              SAfrom=lambda c: c, ReadAlignment=lambda *s: list(s))
    typing_data = eval(body, ns)
    exp = dict(re.findall(r'"(\w+)":\s+"([+.FBG]{9})"', "\n".join(lines(os.path.join(REF, "common/alignments/base.d"), 2616, 2627))))
    with open(os.path.join(HERE, "read_alignment_kat.json"), "w") as f:
        json.dump(dict(cases=typing_data, expect=exp), f)
    # collectReadAlignments KAT  commands/collectPileUps/pileups.d:890-1098: five cases (chains of one read -> read alignments)
    pu = "\n".join(lines(os.path.join(REF, "commands/collectPileUps/pileups.d"), 890, 1098))
    cra = []
    for blk in pu.split("auto alignmentChains = [")[1:]:
        head, tail = blk.split("];", 1)
        chains_ = []
        for m in re.finditer(r"AlignmentChain\(\s*(\d+),\s*Contig\((\d+), (\d+)\),\s*Contig\((\d+), (\d+)\),\s*AlignmentFlags\((\w*)\),\s*"
                             r"\[LocalAlignment\(\s*Locus\((\d+), (\d+)\),\s*Locus\((\d+), (\d+)\),\s*\)\]", head, re.S):
            g = m.groups()
            chains_.append(dict(id=int(g[0]), contigA=[int(g[1]), int(g[2])], contigB=[int(g[3]), int(g[4])], flags=1 if g[5] == "complement" else 0,
                                las=[dict(ab=int(g[6]), ae=int(g[7]), bb=int(g[8]), be=int(g[9]), diffs=0)]))
        expect_ = [[[int(i), sd] for i, sd in re.findall(r"SeededAlignment\(alignmentChains\[(\d+)\], Seed\.(\w+)\)", ra)]
                   for ra in re.findall(r"ReadAlignment\(\[(.*?)\]\)", tail, re.S)]
        cra.append(dict(chains=chains_, expect=expect_))
    with open(os.path.join(HERE, "collect_read_alignments_kat.json"), "w") as f:
        json.dump(cra, f)
    out = dict(source="a-ludi/dentist @ 1aa60e04", tspace=100, ladump=dump, flat=flat, chains=chains, trace_kat=kat)
    with open(os.path.join(HERE, "las_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("flat", len(flat), "chains", len(chains), "dump lines", len(dump), "kat tiles", len(kat["trace"]), "asserts", len(kat["asserts"]),
          "cropper cases", len(cases), "consensus reads", len(cons["reads"]), "collectReadAlignments cases", [(len(c["chains"]), len(c["expect"])) for c in cra], "typing cases", len(typing_data), len(exp), "chain order lists", [len(x) for x in order_kat], "maskcov", len(maskcov["alignments"]), len(maskcov["contigs"]), len(maskcov["mask"]), len(maskcov["changes"]))


if __name__ == "__main__":
    main()
