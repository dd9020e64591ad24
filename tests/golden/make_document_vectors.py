#!/usr/bin/env python
"""Writes tests/golden/document_vectors.json: the byte strings cc.mrlda.Document.write produces
(Document.java:241-263: int32 nEntries, nEntries x (int32 termID, int32 count), int32 nTopics,
nTopics x float64, all big-endian) for the objects the reference's own tests build
(src/test/java/cc/mrlda/DocumentTest.java:120-282): content {1:22, 2:5, 3:10}, gamma
{(double)0.238573f, (double)1.59382f}, and the null-content / null-gamma variants.

Generated with Python's struct module only (independent of the C++ codec under test); HMapII's
iteration order is unspecified, the vectors use ascending term id.
"""
import json
import os
import struct

import numpy as np

CONTENT = [(1, 22), (2, 5), (3, 10)]
GAMMA = [float(np.float32(0.238573)), float(np.float32(1.59382))]


def doc_bytes(content, gamma):
    b = struct.pack(">i", len(content) if content else 0)
    for t, c in content or []:
        b += struct.pack(">ii", t, c)
    b += struct.pack(">i", len(gamma) if gamma else 0)
    for g in gamma or []:
        b += struct.pack(">d", g)
    return b


cases = {
    "content_and_gamma": dict(content=CONTENT, gamma=GAMMA, tokens=37, types=3),
    "content_null_gamma": dict(content=CONTENT, gamma=None, tokens=37, types=3),
    "null_content_gamma": dict(content=None, gamma=GAMMA, tokens=0, types=0),
    "null_content_null_gamma": dict(content=None, gamma=None, tokens=0, types=0),
}
out = {k: dict(v, hex=doc_bytes(v["content"], v["gamma"]).hex()) for k, v in cases.items()}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "document_vectors.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out), "vectors")
