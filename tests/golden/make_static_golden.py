#!/usr/bin/env python
"""Write the static golden fixtures (published known answers + Pillow outputs).

  sha256_kat.json      FIPS 180-4 / NIST CAVP known answers (published constants, typed in
                       here, NOT computed) + padding-boundary lengths.
  fleiss_1971.json     the worked example of Fleiss (1971), 10 subjects x 14 raters x 5
                       categories, kappa = 0.210 (published).
  pillow_resize.npz    small seeded RGB images and the output of Pillow's own
                       Image.resize(..., BILINEAR) for them (Pillow version recorded).

    python tests/golden/make_static_golden.py
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

KAT = [
    {"msg_ascii": "", "repeat": 1,
     "hex": "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"},
    {"msg_ascii": "abc", "repeat": 1,
     "hex": "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"},
    {"msg_ascii": "abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq", "repeat": 1,
     "hex": "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"},
    {"msg_ascii": "abcdefghbcdefghicdefghijdefghijkefghijklfghijklmghijklmn"
                  "hijklmnoijklmnopjklmnopqklmnopqrlmnopqrsmnopqrstnopqrstu", "repeat": 1,
     "hex": "cf5b16a778af8380036ce59e7b0492370b249b11e8f07a51afac45037afee9d1"},
    {"msg_ascii": "a", "repeat": 1000000,
     "hex": "cdc76e5c9914fb9281a1c7e284d73e67f1809a48a497200e046d39ccc7112cd0"},
]
BOUNDARY_LENGTHS = [0, 1, 3, 4, 15, 16, 31, 32, 54, 55, 56, 57, 63, 64, 65, 100, 111, 112,
                    119, 120, 121, 127, 128, 129, 191, 192, 193, 255, 256, 1000, 4095, 4096, 4097]

FLEISS = {
    "source": "Fleiss, J. L. (1971) Measuring nominal scale agreement among many raters, "
              "Psychological Bulletin 76(5) - worked example table",
    "n_raters": 14,
    "table": [[0, 0, 0, 0, 14], [0, 2, 6, 4, 2], [0, 0, 3, 5, 6], [0, 3, 9, 2, 0], [2, 2, 8, 1, 1],
              [7, 7, 0, 0, 0], [3, 2, 6, 3, 0], [2, 5, 3, 2, 2], [6, 5, 2, 1, 0], [0, 2, 2, 3, 7]],
    "P_bar": 0.378, "P_e": 0.213, "kappa": 0.210, "published_decimals": 3,
}

RESIZE_CASES = [  # (in_h, in_w, out_h, out_w)
    (64, 64, 16, 16), (108, 192, 32, 32), (31, 100, 8, 24), (17, 23, 32, 32), (256, 256, 256, 256),
    (270, 480, 64, 64), (5, 7, 16, 16), (1, 1, 4, 4), (300, 257, 256, 256), (512, 512, 256, 256),
    (128, 96, 256, 256), (540, 960, 256, 256),
]


def main():
    with open(os.path.join(HERE, "sha256_kat.json"), "w") as f:
        json.dump({"source": "FIPS 180-4 examples / NIST CAVP SHA-256 ShortMsg+LongMsg",
                   "kat": KAT, "boundary_lengths": BOUNDARY_LENGTHS}, f, indent=1)
    with open(os.path.join(HERE, "fleiss_1971.json"), "w") as f:
        json.dump(FLEISS, f, indent=1)
    import PIL
    from PIL import Image
    rng = np.random.Generator(np.random.Philox(key=[0xB200, 7]))
    out = {"pillow_version": np.array(PIL.__version__)}
    for i, (ih, iw, oh, ow) in enumerate(RESIZE_CASES):
        if ih * iw >= 256 * 256 and (ih, iw) != (300, 257):   # keep the fixture small: smooth images compress well
            yy, xx = np.mgrid[0:ih, 0:iw]
            img = np.stack([(xx // 8 * 37) % 256, (yy // 8 * 91) % 256, ((xx // 16 + yy // 16) * 53) % 256], -1).astype(np.uint8)
        else:
            img = rng.integers(0, 256, size=(ih, iw, 3), dtype=np.uint8)
        thumb = np.asarray(Image.fromarray(img, "RGB").resize((ow, oh), Image.BILINEAR))
        out[f"in_{i}"] = img
        out[f"out_{i}"] = thumb
    np.savez_compressed(os.path.join(HERE, "pillow_resize.npz"), **out)
    print("wrote sha256_kat.json fleiss_1971.json pillow_resize.npz",
          os.path.getsize(os.path.join(HERE, "pillow_resize.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
