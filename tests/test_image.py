"""Image front end, CPU side: the numpy restatement of Pillow's 8-bit bicubic resampler + the reference's tiling + the
CLIP normalisation (oracle/image_oracle.py) against the golden digests minted from the reference's own
`dynamic_preprocess`, Pillow and transformers (oracle/make_golden_image.py) -- bit for bit -- and the host-side
coefficient tables / lookup table of vita_b200.image_frontend against the oracle's."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import image_oracle as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "image_golden.npz"))
CASES = ["wide", "tall", "square", "tiny", "pano"]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_tiles_and_pixels(name):
    img = GOLD[name + "_image"]
    tiles = O.dynamic_preprocess(img)
    assert tiles.shape[0] == int(GOLD[name + "_n_tiles"])
    assert np.array_equal(tiles[0, :32, :32], GOLD[name + "_tile0_crop"])
    assert sha(tiles) == str(GOLD[name + "_tiles_sha256"]), "uint8 tiles differ from Pillow / the reference's tiling"
    px = torch.from_numpy(O.normalize_tiles(tiles)).to(torch.bfloat16)
    assert sha(px.view(torch.int16).numpy()) == str(GOLD[name + "_pixels_bf16_sha256"])


def test_grid_choice_matches_reference_cases():
    # (width, height) -> (columns, rows) as find_closest_aspect_ratio returns them (golden run)
    assert O.closest_grid(500, 300) == (3, 2)
    assert O.closest_grid(700, 1000) == (2, 3)
    assert O.closest_grid(448, 448) == (1, 1)
    assert O.closest_grid(2400, 600) == (4, 1)
    assert O.closest_grid(333, 901) == (2, 5)


def test_frontend_tables_equal_the_oracle():
    from vita_b200 import image_frontend as I
    for in_size, out_size in [(500, 1344), (300, 896), (901, 2240), (37, 896), (3000, 448), (448, 449), (1900, 1792)]:
        ksize, bounds, kk = O.precompute_coeffs(in_size, out_size)
        k2, b2, kk2 = I.resample_tables(in_size, out_size)
        assert k2 == ksize
        assert np.array_equal(b2, bounds.astype(np.int32)) and np.array_equal(kk2, kk.astype(np.int32)), (in_size, out_size)
    for w, h in [(500, 300), (700, 1000), (448, 448), (2400, 600), (333, 901), (61, 37), (950, 120)]:
        assert I.closest_grid(w, h) == O.closest_grid(w, h)
    lut = I.normalize_lut()                         # [3, 256] bf16
    assert np.array_equal(lut.view(torch.int16).numpy(), GOLD["lut_bf16_bits"])
