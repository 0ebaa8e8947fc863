#!/usr/bin/env python3
"""Regenerates the golden fixtures from the reference tree (run in the build
container, where /root/reference exists; the GPU box only sees the committed JSON).

 - ntt_table.json     NTT_TABLE as documented by script/ntt_param.sage:3-132
                      (Falcon vrfy.c GMb table divided by the Montgomery factor 4091)
 - readme_counts.json the constraint-count table of README.md:41-56
 - gadget_kats.json   the known-answer cases of the reference's inline unit tests
                      (gadgets/arithmetics.rs:340-507, gadgets/range_proofs.rs:360-577)
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
Q = 12289


def ntt_table():
    src = open(os.path.join(REF, "script/ntt_param.sage")).read()
    body = src.split("forward = [", 1)[1].split("]", 1)[0]
    raw = [int(x) for x in re.findall(r"\d+", body)]
    assert len(raw) == 1024, len(raw)
    inv = pow(4091, -1, Q)
    return [x * inv % Q for x in raw]


def readme_counts():
    src = open(os.path.join(REF, "README.md")).read()
    rows = {}
    section = None
    for line in src.splitlines():
        if "Falcon-512" in line or "falcon-512" in line:
            section = 512
        if "Falcon-1024" in line or "falcon-1024" in line:
            section = 1024
        m = re.match(r"\|?\s*([A-Za-z ]+?)\s*\|\s*(\d+)\s*\|\s*(\d+)\s*\|\s*(\d+)\s*\|", line)
        if m and section:
            rows["%d/%s" % (section, m.group(1).strip())] = [int(m.group(i)) for i in (2, 3, 4)]
    return rows


def gadget_kats():
    # transcribed from the reference's test macros (value, expected, satisfied)
    M = Q
    B512, B1024 = 34034726, 70265242
    return {
        # arithmetics.rs:346-361
        "mod_q": [[6, 6, True], [0, 0, True], [M, 0, True], [M + 1, 1, True], [6, 7, False], [5, M - 1, False]],
        # arithmetics.rs:418-432
        "mul_mod": [[6, 7, 42, True], [0, 100, 0, True], [100, 0, 0, True], [5, 12288, 12284, True],
                    [6, 7, 41, False], [5, 12288, 12283, False]],
        # arithmetics.rs:480-494
        "add_mod": [[6, 36, 42, True], [0, 100, 100, True], [100, 0, 100, True], [5, M - 1, 4, True],
                    [6, 7, 41, False], [5, M - 1, 3, False]],
        # range_proofs.rs:365-389
        "less_than_q": [[42, True], [0, True], [1 << 12, True], [1 << 13, True], [M - 1, True], [M, False],
                        [M + 1, False], [M * 10000, False]],
        # range_proofs.rs:529-547
        "less_than_6144": [[42, True], [0, True], [6143, True], [6144, False], [6145, False], [M, False]],
        # range_proofs.rs:442-474 (cfg falcon-512 / falcon-1024)
        "norm_bound_512": [[42, True], [0, True], [1 << 25, True], [1 << 24, True], [B512 - 1, True],
                           [B512, False], [B512 + 1, False], [1 << 26, False], [1 << 27, False]],
        "norm_bound_1024": [[42, True], [0, True], [1 << 25, True], [1 << 24, True], [1 << 26, True],
                            [B1024 - 1, True], [B1024, False], [B1024 + 1, False], [1 << 27, False]],
    }


def main():
    tab = ntt_table()
    assert all(tab[i] == pow(7, int(format(i, "010b")[::-1], 2), Q) for i in range(1024))
    json.dump(tab, open(os.path.join(HERE, "ntt_table.json"), "w"))
    json.dump(readme_counts(), open(os.path.join(HERE, "readme_counts.json"), "w"), indent=1)
    json.dump(gadget_kats(), open(os.path.join(HERE, "gadget_kats.json"), "w"), indent=1)
    print("ok", readme_counts())


if __name__ == "__main__":
    main()
