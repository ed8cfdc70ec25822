"""Close the rasterization pin the day a box with the reference's dependency stack exists.

Needs the reference's pins: rasterio 1.3.2 on GDAL 3.0.4, affine 2.3.1 (requirements.txt:9,77,211) -- NOT importable in the build
container of this repository (no network, no GDAL), which is why DESIGN.md calls the rasterization oracle "parity unpinned".
Run there:

    python tests/golden/make_rasterio_golden.py            # writes tests/golden/rasterio_masks.npz

and commit the file: tests/test_rasterio_golden.py then holds oracle/gdal_fill.py, the C oracle and (with -m gpu) the CUDA
kernels to what rasterio itself returns for
  * rasterio.features.rasterize(shapes, out_shape, transform)            (scripts/sandbox/add_tile_mask.py:112-113)
  * rasterio.mask.mask(dataset, shapes, crop=True)                        (scripts/functions/fct_misc.py:77)
on the hand-derived known-answer shapes of tests/test_oracle_kat.py plus seeded random and adversarial polygons (the same
generator as tests/test_gpu_properties.py).  The polygons are stored with the outputs, so the test needs nothing else.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def polygons():
    """(name, rings) cases: KAT shapes, adversarial lattice polygons, random multi-ring polygons"""
    from test_gpu_properties import _adversarial_polygons
    from test_oracle_kat import ring
    cases = [
        ("integer_rect", [ring((2, 1), (6, 1), (6, 4), (2, 4))]),
        ("on_centres_cw", [ring((1.5, 1.5), (4.5, 1.5), (4.5, 3.5), (1.5, 3.5))]),
        ("on_centres_ccw", [ring((1.5, 1.5), (1.5, 3.5), (4.5, 3.5), (4.5, 1.5))]),
        ("triangle_vertex_on_scanline", [ring((1, 0), (7, 0), (4, 2.5))]),
        ("hole", [ring((0, 0), (10, 0), (10, 10), (0, 10)), ring((3, 3), (3, 7), (7, 7), (7, 3))]),
        ("even_odd_parts", [ring((0, 0), (6, 0), (6, 4), (0, 4)), ring((4, 0), (10, 0), (10, 4), (4, 4))]),
    ]
    for i, rings in enumerate(_adversarial_polygons(3000, seed=424242)):
        cases.append((f"adv{i}", rings))
    return cases


TRANSFORMS = [(1.0, 0.0, 0.0, 0.0, 1.0, 0.0), (0.5, 0.0, 1000.0, 0.0, -0.5, 5000.0),
              (0.5971642834779395, 0.0, 829045.2, 0.0, -0.5971642834779395, 5933729.9)]
W, H = 32, 24


def main():
    import rasterio
    from affine import Affine
    from rasterio.features import rasterize
    from rasterio.io import MemoryFile
    from rasterio.mask import mask as rio_mask
    print("rasterio", rasterio.__version__, "GDAL", rasterio.__gdal_version__)
    names, xy, ring_off, poly_off, tr_idx = [], [], [0], [0], []
    full_masks, crop_masks, crop_windows = [], [], []
    data = (np.arange(H * W, dtype=np.int32).reshape(1, H, W) % 251 + 1).astype(np.uint8)
    for name, rings in polygons():
        for ti, t in enumerate(TRANSFORMS):
            aff = Affine(*t)
            world = [np.stack([t[2] + r[:, 0] * t[0], t[5] + r[:, 1] * t[4]], 1) for r in rings]
            geom = {"type": "Polygon", "coordinates": [w.tolist() for w in world]}       # first ring exterior, others holes
            m = rasterize([(geom, 1)], out_shape=(H, W), transform=aff, fill=0, dtype="uint8")
            with MemoryFile() as mf:
                with mf.open(driver="GTiff", height=H, width=W, count=1, dtype="uint8", transform=aff) as ds:
                    ds.write(data)
                    try:
                        out, out_t = rio_mask(ds, [geom], crop=True)
                        win = (int(round((out_t.c - aff.c) / aff.a)), int(round((out_t.f - aff.f) / aff.e)), out.shape[2], out.shape[1])
                        cm = np.zeros((H, W), np.uint8)
                        cm[win[1]:win[1] + win[3], win[0]:win[0] + win[2]] = out[0] != 0
                    except ValueError:                                                   # "Input shapes do not overlap raster."
                        win, cm = (-1, -1, 0, 0), np.zeros((H, W), np.uint8)
            names.append(name)
            tr_idx.append(ti)
            for r in world:
                xy.append(r)
                ring_off.append(ring_off[-1] + len(r))
            poly_off.append(len(ring_off) - 1)
            full_masks.append(np.packbits(m))
            crop_masks.append(np.packbits(cm))
            crop_windows.append(win)
    out = os.path.join(HERE, "rasterio_masks.npz")
    np.savez_compressed(out, names=np.array(names), xy=np.concatenate(xy), ring_off=np.array(ring_off), poly_off=np.array(poly_off),
                        transform_index=np.array(tr_idx), transforms=np.array(TRANSFORMS), shape=np.array([H, W]),
                        rasterize=np.stack(full_masks), mask_crop=np.stack(crop_masks), crop_window=np.array(crop_windows),
                        versions=np.array([rasterio.__version__, rasterio.__gdal_version__]))
    print("wrote", out, os.path.getsize(out), "bytes,", len(names), "cases")


if __name__ == "__main__":
    main()
