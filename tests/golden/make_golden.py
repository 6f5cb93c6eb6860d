#!/usr/bin/env python3
"""Generates the golden fixtures in this directory FROM THE REFERENCE SOURCES (run in the authoring
container where /root/reference exists; the fixtures travel, the reference does not).

  prn1_ca_code.json   the 1023-chip PRN-1 known-answer vector of src/bk/gps_ca_prn.rs:72-123
  ca_table.json       sha256 of the 32x1023 i8 table src/constants/gps_ca_constants.rs (row-major bytes)
                      plus the first 16 chips of every row
  config_txt.json     the PRN / carrier / code-phase table of src/test_data/GPS_recordings/config.txt:8-17
"""
import hashlib
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def ints(text):
    return [int(v) for v in re.findall(r"-?\d+", text)]


def main():
    src = open(os.path.join(REF, "src/bk/gps_ca_prn.rs")).read()
    body = src[src.index("ca_code_prn1,"):]
    body = body[body.index("vec!["):body.index("]")]
    prn1 = ints(body.replace("vec![", ""))
    assert len(prn1) == 1023 and set(prn1) == {1, -1}
    json.dump({"source": "src/bk/gps_ca_prn.rs:72-123", "chips": prn1}, open(os.path.join(HERE, "prn1_ca_code.json"), "w"))

    tab = open(os.path.join(REF, "src/constants/gps_ca_constants.rs")).read()
    tab = tab[tab.index("= [") + 3:]
    vals = ints(tab)
    assert len(vals) == 32 * 1023
    b = bytes((v & 0xFF) for v in vals)
    json.dump({"source": "src/constants/gps_ca_constants.rs", "sha256": hashlib.sha256(b).hexdigest(),
               "first16": [vals[r * 1023:r * 1023 + 16] for r in range(32)]},
              open(os.path.join(HERE, "ca_table.json"), "w"))

    cfg = open(os.path.join(REF, "src/test_data/GPS_recordings/config.txt")).read()
    rows = []
    for line in cfg.splitlines():
        m = re.match(r"\s*(\d+)(?:\[\d\])?\s+(\d\.\d+)\s+(\d+)\s*$", line)
        if m:
            rows.append({"prn": int(m.group(1)), "carrier_mhz": float(m.group(2)), "code_phase": int(m.group(3))})
    assert len(rows) == 10
    json.dump({"source": "src/test_data/GPS_recordings/config.txt:2-17", "fs": 16367600, "if": 4130400, "rows": rows},
              open(os.path.join(HERE, "config_txt.json"), "w"))
    print("wrote golden fixtures")


if __name__ == "__main__":
    main()
