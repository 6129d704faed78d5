"""Generate tests/golden/*.json from the reference tree and from the oracle.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Outputs (committed):
  ref_tables.json      literal tables parsed out of the reference's headers: ZZ, YQuantumTb, CQuantumTb
                       (src/jpezy.hpp), the (size, code) LUTs and the four DHT byte arrays
                       (src/encoder/huffman_table.hpp)
  oracle_vectors.json  SHA-256 of oracle outputs on the seeded synthetic images (regression pins)
  ref_vectors.json     outputs of the REFERENCE'S OWN encoder / decoder classes (oracle/_ref: /root/reference/src compiled
                       unmodified against oracle/shim) on seeded synthetic images: file length and SHA-256, wrote_size(),
                       SHA-256 of the decoded planes (colour and --gray), and for the small cases the file itself (hex)
"""
import hashlib
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/src"


def ints(body):
    body = re.sub(r"//.*", "", body)
    return [int(t.replace("_byte", ""), 0) for t in re.findall(r"0x[0-9a-fA-F]+(?:_byte)?|\d+", body)]


def array(text, name):
    m = re.search(re.escape(name) + r"\s*\{(.*?)\};", text, re.S)
    assert m, name
    return ints(m.group(1))


def main():
    jp = open(os.path.join(REF, "jpezy.hpp")).read()
    ht = open(os.path.join(REF, "encoder", "huffman_table.hpp")).read()
    out = {"ZZ": array(jp, "ZZ"), "YQuantumTb": array(jp, "YQuantumTb"), "CQuantumTb": array(jp, "CQuantumTb")}
    for n in ["YDcSizeT", "YDcCodeT", "CDcSizeT", "CDcCodeT", "YAcSizeT", "YAcCodeT", "CAcSizeT", "CAcCodeT", "YDcDht", "CDcDht",
              "YAcDht", "CAcDht"]:
        out[n] = array(ht, n)
    assert len(out["ZZ"]) == 64 and len(out["YAcSizeT"]) == 162 and len(out["YAcDht"]) == 183 and len(out["YDcDht"]) == 33
    json.dump(out, open(os.path.join(HERE, "ref_tables.json"), "w"), indent=0)

    import numpy as np
    import jpezy_b200 as J
    import oracle as orc
    o = orc.Oracle()
    vec = {}
    for name, fam, W, H, gray in [("c1_photo_512", 0, 512, 512, False), ("c1_noise_512", 1, 512, 512, False),
                                  ("adversarial_200x120", 2, 200, 120, False), ("photo_gray_333x77", 0, 333, 77, True),
                                  ("tiny_1x1", 0, 1, 1, False)]:
        r, g, b = J.synth.image(fam, W, H)
        f = o.encode(r, g, b, W, H, gray=gray)
        c = o.coefs(r, g, b, W, H, gray=gray)
        _, _, R, G, B = o.decode(f, gray=gray)
        vec[name] = {"family": fam, "W": W, "H": H, "gray": gray, "file_bytes": len(f),
                     "input_sha256": hashlib.sha256(r.tobytes() + g.tobytes() + b.tobytes()).hexdigest(),
                     "file_sha256": hashlib.sha256(f).hexdigest(),
                     "coefs_sha256": hashlib.sha256(np.ascontiguousarray(c).tobytes()).hexdigest(),
                     "decoded_sha256": hashlib.sha256(R.tobytes() + G.tobytes() + B.tobytes()).hexdigest()}
    json.dump(vec, open(os.path.join(HERE, "oracle_vectors.json"), "w"), indent=1)
    ref = orc.Reference()
    rv = {}
    cases = [("photo_64x48", 0, 64, 48, False), ("noise_200x120", 1, 200, 120, False), ("adversarial_136x72", 2, 136, 72, False),
             ("ragged_17x33", 0, 17, 33, False), ("ragged_250x7", 1, 250, 7, False), ("tiny_1x1", 0, 1, 1, False),
             ("one_mcu_16x16", 2, 16, 16, False), ("gray_333x77", 0, 333, 77, True), ("gray_noise_40x25", 1, 40, 25, True),
             ("c1_photo_512", 0, 512, 512, False), ("c1_noise_512", 1, 512, 512, False), ("hd_band_1920x64", 0, 1920, 64, False)]
    for name, fam, W, H, gray in cases:
        r, g, b = J.synth.image(fam, W, H)
        f, wrote = ref.encode(r, g, b, W, H, gray=gray)
        e = {"family": fam, "W": W, "H": H, "gray": gray, "file_bytes": len(f), "wrote_size": wrote,
             "input_sha256": hashlib.sha256(r.tobytes() + g.tobytes() + b.tobytes()).hexdigest(),
             "file_sha256": hashlib.sha256(f).hexdigest()}
        for mode, gm in (("decoded_sha256", False), ("decoded_gray_sha256", True)):
            Wd, Hd, R, G, B = ref.decode(f, gray=gm)
            assert (Wd, Hd) == (W, H)
            e[mode] = hashlib.sha256(R.tobytes() + G.tobytes() + B.tobytes()).hexdigest()
        e["plane_bytes"] = int(R.size)
        if len(f) <= 1200:
            e["file_hex"] = f.hex()
        rv[name] = e
    cos, ds = ref.constants()
    rv["_constants"] = {"cos_table_hex": [float(c).hex() for c in cos], "dis_sqrt_hex": float(ds).hex()}
    json.dump(rv, open(os.path.join(HERE, "ref_vectors.json"), "w"), indent=1)
    print("wrote ref_tables.json, oracle_vectors.json, ref_vectors.json")


if __name__ == "__main__":
    main()
