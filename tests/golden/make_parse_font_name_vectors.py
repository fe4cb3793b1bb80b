#!/usr/bin/env python
"""Extracts the known-answer vectors of the reference's own test `test_parse_font_name`
(reference src/font/parse_font_name.rs, the `samples` list: family;postscript;family;style;weight;width)
into tests/golden/parse_font_name_vectors.json.  Run in the build container (needs /root/reference)."""
import json
import os
import re

SRC = "/root/reference/src/font/parse_font_name.rs"
text = open(SRC).read()
body = text[text.index("fn test_parse_font_name"):]
rows = re.findall(r'"([^"\n]*;[^"\n]*;[^"\n]*;[^"\n]*;\d+;[^"\n]*)"', body)
vectors = []
for r in rows:
    fam, ps, efam, style, weight, width = r.split(";")
    vectors.append({"family": fam, "ps_name": ps, "expect": [efam, style, int(weight), width]})
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "parse_font_name_vectors.json")
json.dump({"source": "reference src/font/parse_font_name.rs test_parse_font_name", "vectors": vectors}, open(out, "w"), indent=0)
print(len(vectors), "vectors ->", out)
