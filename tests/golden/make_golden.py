"""Extract the golden vectors the reference's notebooks print into JSON fixtures.

Run in the build container only (reads /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

Sources (raw notebook outputs; TF 1.12.0 per antisymmetric_conv_kernel.ipynb cell 0):
  * antisymmetric_conv_kernel.ipynb cells 1-3: tf.nn.conv2d known-answer vector
    (7x7x1 image, 3x3x1x1 kernel, NHWC / SAME / stride 1).
  * experiments_antisymmetric_resnet_v6.ipynb cell 26: get_kernel() of a trained
    64-channel antisymmetric layer: K[:,:,10,31], K[:,:,31,10], K[:,:,4,4].
  * experiments_antisymmetric_resnet_v6.ipynb cell 41: integer NumPy prototype
    of the assembly loop (512 channels): six printed 3x3 blocks.
  * experiments_antisymmetric_resnet_v6.ipynb cell 35: 7x7 anti-centrosymmetric
    integer matrix printed by the general-k prototype.
  * antisymmetric_conv_kernel.ipynb / experiments_antisymmetric_resnet_v2.0.ipynb:
    printed diagonal blocks K[:,:,1,1] of trained antisymmetric layers.
"""
import json
import os
import re

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def cell_text(nb, i):
    out = []
    for o in nb["cells"][i].get("outputs", []):
        if "text" in o:
            out.append("".join(o["text"]))
        elif "data" in o and "text/plain" in o["data"]:
            out.append("".join(o["data"]["text/plain"]))
    return "\n".join(out)


def numbers(text):
    return [float(t) for t in re.findall(r"-?\d+\.\d*(?:e-?\d+)?|-?\d+", text)]


def blocks(text):
    """Split printed output into blank-line separated groups of numbers."""
    res = []
    for blk in re.split(r"\n\s*\n", text.strip()):
        n = numbers(blk)
        if n:
            res.append(n)
    return res


def main():
    nb = json.load(open(os.path.join(REF, "antisymmetric_conv_kernel.ipynb")))
    img = numbers(cell_text(nb, 1).split("shape=")[0])
    ker = numbers(cell_text(nb, 2).split("numpy=")[1].split("dtype")[0])
    out = numbers(cell_text(nb, 3).split("shape=")[0])
    assert len(img) == 49 and len(ker) == 9 and len(out) == 49
    json.dump({"source": "antisymmetric_conv_kernel.ipynb cells 1-3 (TF 1.12.0 tf.nn.conv2d, NHWC, SAME, stride 1)",
               "image_7x7": img, "kernel_3x3": ker, "output_7x7": out},
              open(os.path.join(HERE, "conv2d_known_answer.json"), "w"), indent=1)

    nb6 = json.load(open(os.path.join(REF, "experiments_antisymmetric_resnet_v6.ipynb")))
    b = blocks(cell_text(nb6, 26))
    assert b[0] == [3, 3, 64, 64] and all(len(x) == 9 for x in b[1:4])
    json.dump({"source": "experiments_antisymmetric_resnet_v6.ipynb cell 26 (get_kernel() of a trained 64-ch layer)",
               "shape": b[0], "K_10_31": b[1], "K_31_10": b[2], "K_4_4": b[3]},
              open(os.path.join(HERE, "kernel_structure_v6_cell26.json"), "w"), indent=1)

    b = blocks(cell_text(nb6, 41))
    assert len(b) == 6 and all(len(x) == 9 for x in b)
    json.dump({"source": "experiments_antisymmetric_resnet_v6.ipynb cell 41 (integer prototype, 512 channels)",
               "K_0_0": b[0], "K_1_0": b[1], "K_0_1": b[2], "K_100_100": b[3], "K_14_234": b[4], "K_234_14": b[5]},
              open(os.path.join(HERE, "kernel_structure_v6_cell41.json"), "w"), indent=1)

    m = numbers(cell_text(nb6, 35))
    assert len(m) == 49
    json.dump({"source": "experiments_antisymmetric_resnet_v6.ipynb cell 35 (7x7 anti-centrosymmetric prototype)",
               "matrix_7x7": m},
              open(os.path.join(HERE, "centrosymmetric_7x7_v6_cell35.json"), "w"), indent=1)
    # Diagonal blocks K[:,:,1,1] of trained antisymmetric layers printed by two notebooks (gamma = 0): every cell that
    # prints `res2a_branch2_kernel[:,:,1,1]` / `res2d_branch2_kernel[:,:,1,1]`.
    diag = []
    for fn in ("antisymmetric_conv_kernel.ipynb", "experiments_antisymmetric_resnet_v2.0.ipynb"):
        nbx = json.load(open(os.path.join(REF, fn)))
        for i, c in enumerate(nbx["cells"]):
            src = "".join(c.get("source", []))
            if c.get("cell_type") == "code" and "res2a_branch2_kernel[:,:,1,1]" in src and "res2d_branch2_kernel[:,:,1,1]" in src:
                bl = blocks(cell_text(nbx, i))
                if len(bl) == 2 and all(len(x) == 9 for x in bl):
                    diag.append({"source": "%s cell %d" % (fn, i), "res2a_K_1_1": bl[0], "res2d_K_1_1": bl[1]})
    assert len(diag) >= 2
    json.dump({"source": "printed get_kernel()[:,:,1,1] / layer.kernel[:,:,1,1] of trained Conv2DAntisymmetric layers (gamma = 0)",
               "cells": diag}, open(os.path.join(HERE, "diagonal_blocks.json"), "w"), indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
