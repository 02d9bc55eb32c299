"""Golden vectors for the tiler: runs the UNMODIFIED source of the reference's ``calculate_slice_bboxes``
(restoration/utils.py:332-375).  utils.py cannot be imported here (astropy, photutils, sep, ndpatch, reproject are
absent), so the function's own AST node is compiled and executed on its own — nothing of it is copied into the repo;
only its outputs are committed (tiles_golden.json).  Run in the build container: python tests/golden/make_tiles_golden.py
"""
import ast
import json
import os

SRC = "/root/reference/restoration/utils.py"
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [  # (H, W, sh, sw, ratio_h, ratio_w)
    (2048, 2048, 256, 256, 0.0, 0.0),            # BASELINE config 4: 8 x 8 tiles
    (2048, 2048, 256, 256, 10 / 256, 10 / 256),  # any overlap -> 9 x 9 = 81 tiles
    (1978, 1978, 256, 256, 10 / 256, 10 / 256),  # 64 overlapping tiles
    (375, 375, 100, 100, 10 / 100, 10 / 100),    # create_subdivisions defaults on the paper's sub-frame
    (450, 375, 128, 64, 0.2, 0.3),
    (512, 512, 512, 512, 0.2, 0.2),              # the function's own defaults, one tile
    (100, 300, 256, 256, 0.2, 0.2),              # image smaller than the tile in one direction
    (31, 31, 32, 32, 0.0, 0.0),                  # smaller in both
    (1000, 777, 100, 90, 7 / 100, 7 / 90),
    (640, 480, 64, 64, 0.5, 0.25),
]


def reference_function():
    tree = ast.parse(open(SRC).read())
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "calculate_slice_bboxes")
    ns = {}
    exec(compile(ast.Module(body=[node], type_ignores=[]), SRC, "exec"), ns)
    return ns["calculate_slice_bboxes"]


if __name__ == "__main__":
    f = reference_function()
    out = [{"args": list(c), "boxes": f(*c)} for c in CASES]
    json.dump(out, open(os.path.join(HERE, "tiles_golden.json"), "w"))
    print("wrote", len(out), "cases;", [len(o["boxes"]) for o in out], "boxes")
