"""Generates tests/golden/ref_outputs_v1.npz: outputs of the reference's OWN sources for the pieces round 1 left unpinned.

  * SVD<Power> and PM (include/SVD_class.hpp:184-219, src/PM.cpp:4-81)            -- oracle/_ref/libref_rsvd.so
  * the OLDER API: 5-argument rSVD, intermediate_step, singularValueDecomposition, powerMethod, the QR class
    (image_compression/src/*.cpp)                                                    -- oracle/_ref/libref_imgcomp.so
  * the Image class (image_compression/src/image_com.cpp) on a PGM file, through the stb headers the reference vendors

Both libraries are the reference's first-party translation units compiled from /root/reference by oracle/Makefile.  The power
method starts from std::random_device and the older rSVD draws Omega the same way, so these are TOLERANCE pins on inputs whose
answers do not depend on the start (geometric spectra, exactly low-rank matrices).  Run in the dev container:

    python tests/golden/make_golden_v1.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import rsvd_oracle as O  # noqa: E402


def geometric(m, n, r, ratio, seed):
    """U diag(ratio^j) V^T with orthonormal factors: singular values known exactly, gaps fixed by `ratio`."""
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((m, r))); V, _ = np.linalg.qr(rng.standard_normal((n, r)))
    return np.asfortranarray((U * ratio ** np.arange(r)) @ V.T)


def power_inputs():
    return {"geo_12x40_r12": geometric(12, 40, 12, 0.5, 1), "geo_30x20_r8": geometric(30, 20, 8, 0.4, 2), "geo_16x100_r16": geometric(16, 100, 16, 0.5, 3)}


def v1_inputs():
    """name -> (A, l): exactly rank-r inputs with r <= l - 2, so the older rSVD's answer does not depend on its random Omega."""
    return {"rank8_120x90_l15": (geometric(120, 90, 8, 0.5, 11), 15), "rank6_64x200_l16": (geometric(64, 200, 6, 0.4, 12), 16),
            "rank10_300x40_l20": (geometric(300, 40, 10, 0.6, 13), 20)}


def qr_inputs():
    rng = np.random.default_rng(21)
    return {"qr_test2_4x3": np.asfortranarray(np.arange(1.0, 13.0).reshape(4, 3)), "gauss_40x12": np.asfortranarray(rng.standard_normal((40, 12))),
            "gauss_25x25": np.asfortranarray(rng.standard_normal((25, 25)))}


def image_pixels(height=96, width=128, seed=31):
    """An 8-bit grey picture (height x width, row-major like a decoded file): smooth low-rank content plus a little texture."""
    rng = np.random.default_rng(seed)
    y = np.arange(height)[:, None] / height; x = np.arange(width)[None, :] / width
    img = 120 + 70 * np.sin(2 * np.pi * y) * np.cos(3 * np.pi * x) + 40 * (x - 0.5) * (y - 0.3) + 6 * rng.standard_normal((height, width))
    return np.clip(np.round(img), 0, 255).astype(np.uint8)


def write_pgm(path, pixels):
    h, w = pixels.shape
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (w, h)); f.write(pixels.tobytes())


IMAGE_CASES = [(1, 12, (96, 128)), (2, 8, (96, 96))]          # (downscale factor, k, (height, width) of the picture)


def main():
    ref = O.RefLib(); ref1 = O.RefLibV1()
    out = {}
    for nm, B in power_inputs().items():
        U, S, V = ref.svd(B, O.POWER)                                      # SVD<Power>(B).compute()
        out[f"svdpower/{nm}/S"] = S; out[f"svdpower/{nm}/shapes"] = np.array(list(U.shape) + list(V.shape))
        k = len(S)
        out[f"svdpower/{nm}/absU"] = np.abs(U[:, :k]); out[f"svdpower/{nm}/absVrows"] = np.abs(V[:k, :])
        U3, S3, V3 = ref.svd(B, O.POWER, r=3)
        out[f"svdpower/{nm}/r3/S"] = S3; out[f"svdpower/{nm}/r3/shapes"] = np.array(list(U3.shape) + list(V3.shape))
        sg, u, v = ref.pm(B)                                               # PM(A, B, sigma, u, v)
        out[f"pm/{nm}/sigma"] = np.array(sg); out[f"pm/{nm}/absu"] = np.abs(u); out[f"pm/{nm}/absv"] = np.abs(v)
        sg1, u1, v1 = ref1.power_method(B)                                 # older powerMethod
        out[f"v1/pm/{nm}/sigma"] = np.array(sg1); out[f"v1/pm/{nm}/absu"] = np.abs(u1); out[f"v1/pm/{nm}/absv"] = np.abs(v1)
        dim = min(6, min(B.shape))
        U2, S2, V2 = ref1.svd(B, dim)                                      # older singularValueDecomposition
        out[f"v1/svd/{nm}/S"] = S2; out[f"v1/svd/{nm}/absU"] = np.abs(U2); out[f"v1/svd/{nm}/absV"] = np.abs(V2)
    rng = np.random.default_rng(5)
    for nm, (A, l) in v1_inputs().items():
        Om = np.asfortranarray(rng.standard_normal((A.shape[1], l)))
        out[f"v1/istep/{nm}/Omega"] = Om
        out[f"v1/istep/{nm}/Q"] = ref1.intermediate_step(A, Om, l, 1)      # Givens-QR range finder, q = 1
        U, S, V = ref1.rsvd(A, l)                                          # Omega from std::random_device inside
        out[f"v1/rsvd/{nm}/S"] = S; out[f"v1/rsvd/{nm}/shapes"] = np.array(list(U.shape) + list(V.shape))
        out[f"v1/rsvd/{nm}/err"] = np.array(np.linalg.norm(A - (U * S) @ V.T))
    for nm, A in qr_inputs().items():
        for red in (1, 0):
            Q, R = ref1.qr(A, bool(red))                                   # QRReducedDecomposition / QRFullDecomposition
            out[f"v1/qr/{nm}/red{red}/absQ"] = np.abs(Q); out[f"v1/qr/{nm}/red{red}/R"] = R
    # Image: load (stb, PGM) -> [downscale] -> normalize -> compress(k) -> reconstruct
    # (downscale / upscale index the transposed matrix with (row, column) swapped -- image_com.cpp:199-203 -- so they are only well
    # defined for square pictures: the downscale case uses a square crop)
    tmp = Path("/tmp/rsvdb_golden_v1.pgm")
    for scale, k, (h, w) in IMAGE_CASES:
        px = image_pixels()[:h, :w]; write_pgm(tmp, px)
        d = ref1.image_flow(tmp, w // scale, h // scale, scale, k)
        key = f"v1/image/s{scale}_k{k}/"
        out[key + "norm"] = d["norm"]; out[key + "range"] = np.array([d["lo"], d["hi"]]); out[key + "S"] = d["S"]
        out[key + "recon_err"] = np.array(np.linalg.norm(d["norm"] - d["recon"])); out[key + "ratio"] = np.array(d["ratio"])
    np.savez_compressed(Path(__file__).with_name("ref_outputs_v1.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
