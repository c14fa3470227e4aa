"""Regenerates tests/golden/dodecahedron_bih.json from the reference's own tree dump.

Source: /root/reference/BIH_Raytracer/BIH_Raytracer/BIH1.txt (identical to BIH2.txt and to each of
the 74 frames in log.txt), printed by the commented block R/src/Renderer.cpp:617-636.  This is the
only pinned result in the reference; the mesh that produced it is reconstructed in SURVEY.md
Appendix A (bihrt.scenes.dodecahedron).  Run in the dev container (the reference tree does not
exist on the GPU box):  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import re

REF = "/root/reference/BIH_Raytracer/BIH_Raytracer"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dodecahedron_bih.json")


def parse_dump(text):
    nodes = []
    for blk in re.split(r"\n\s*\n", text.strip()):
        lines = [l.strip() for l in blk.strip().splitlines() if l.strip()]
        if not lines or not lines[0].startswith("NODE"):
            continue
        d = {"node": int(lines[0].split()[1])}
        for l in lines[1:]:
            k, v = l.split(":")
            v = v.strip()
            d[k.strip()] = {"TRUE": True, "FALSE": False}.get(v, v)
        nodes.append({
            "node": d["node"], "parent": int(d["parent"]),
            "children": [int(d["leftChild"]), int(d["rightChild"])], "axis": int(d["axis"]),
            "is_leaf": [bool(d["isLeftLeaf"]), bool(d["isRightLeaf"])],
            "clip": [float(d["clipPlaneLEFT"]), float(d["clipPlaneRIGHT"])]})
    return nodes


def main():
    texts = {n: open(os.path.join(REF, n)).read() for n in ("BIH1.txt", "BIH2.txt")}
    nodes = parse_dump(texts["BIH1.txt"])
    assert nodes == parse_dump(texts["BIH2.txt"])
    log = open(os.path.join(REF, "log.txt")).read()
    allnodes = parse_dump(log)
    frames = [allnodes[i:i + len(nodes)] for i in range(0, len(allnodes), len(nodes))]
    same = sum(f == nodes for f in frames)
    out = {"source": "R/BIH1.txt == R/BIH2.txt == %d/%d frames of R/log.txt" % (same, len(frames)),
           "sha256_BIH1": hashlib.sha256(texts["BIH1.txt"].encode()).hexdigest(),
           "print_precision": "operator<<(float), 6 significant digits",
           "nodes": nodes}
    json.dump(out, open(OUT, "w"), indent=1)
    print("wrote", OUT, len(nodes), "nodes;", out["source"])


if __name__ == "__main__":
    main()
