#!/usr/bin/env python
"""Generates tests/golden/binary_*.{npz,json} from the reference's OWN sources run on the CPU.

Run in the build container only (needs /root/reference):  python tests/golden/gen_binary_golden.py
It builds oracle/_ref/libbldpc_ref_*.so with oracle/build_ref.sh (CPU shim, see that script),
then records
  * sha256 of H / Wc / Wv / Address_Variablenode (B/Simulation.cu:292-387) for several H files,
  * one 16-frame batch for J4_L24_Z96: channel samples from AWGNChannel_CPU (seeds 173,173,173,
    Es/N0 3 dB), and LDPC_Decoder_GPU's output D (+ iteration count) for the literal and the
    one-line-fixed Transform_H,
  * Simulation_GPU counters (frames, error frames, error bits, iterations, false, alarm) at
    several Es/N0 points, F=256, maxIT 10, leastTestFrames 1024.
"""
import ctypes as C, hashlib, json, os, subprocess, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "tests", "golden")
ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))


def build(name, J, L, Z, hfile, F, maxit, least, snrtype, variant):
    subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh"), name, str(J), str(L), str(Z), hfile,
                           str(F), str(maxit), str(least), str(snrtype), variant])
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", f"libbldpc_ref_{name}.so"))
    lib.ref_init()
    lib.ref_sim_point.argtypes = [C.c_float, C.POINTER(C.c_long)]
    return lib


def tables(lib):
    g = np.zeros(10, np.int32)
    lib.ref_geometry(ip(g))
    J, L, Z, N, K, M, F, maxit, Wcm, Wvm = map(int, g)
    H = np.zeros(J * L, np.int32); Wc = np.zeros(J + 1, np.int32); Wv = np.zeros(L + 1, np.int32)
    addr = np.zeros(N * Wvm, np.int32)
    lib.ref_tables(ip(H), ip(Wc), ip(Wv), ip(addr))
    return g, H, Wc, Wv, addr


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    meta = {"tables": {}, "sim_points": []}
    codes = [("J4_L24_Z96_BlockH.txt", 4, 24, 96), ("J15_L30_Z1280_BlockH.txt", 15, 30, 1280),
             ("PON_LDPC.txt", 12, 69, 256), ("J10_L60_Z160_BlockH.txt", 10, 60, 160),
             ("J32_L64_Z64_BlockH.txt", 32, 64, 64), ("J48_L60_Z160_BlockH.txt", 48, 60, 160)]
    for hfile, J, L, Z in codes:
        for variant in ("literal", "fixed"):
            tag = f"tab_J{J}L{L}Z{Z}_{variant}"  # unique: dlopen caches by path
            lib = build(tag, J, L, Z, hfile, 16, 10, 1024, 1, variant)
            g, H, Wc, Wv, addr = tables(lib)
            meta["tables"].setdefault(hfile, {"J": J, "L": L, "Z": Z, "H": H.tolist(), "Wc": Wc.tolist(),
                                              "Wv": Wv.tolist()})[f"addr_sha256_{variant}"] = sha(addr)
            os.remove(os.path.join(ROOT, "oracle", "_ref", f"libbldpc_ref_{tag}.so"))
    # one 16-frame batch, J4_L24_Z96, Es/N0 = 3 dB
    batch = {}
    for variant in ("literal", "fixed"):
        lib = build(f"C1_{variant}_F16", 4, 24, 96, "J4_L24_Z96_BlockH.txt", 16, 10, 1024, 1, variant)
        g, *_ = tables(lib)
        N, F = int(g[3]), int(g[6])
        seed = np.array([173, 173, 173], np.int32)
        sigma = np.float32(np.sqrt(0.5 / 10 ** 0.3))
        y = np.zeros(N * F, np.float32)
        lib.ref_awgn(ip(seed), C.c_float(float(sigma)), fp(y))
        D = np.zeros((N + 1) * F, np.int32)
        it = lib.ref_decode(fp(y), ip(D))
        batch["y"] = y.reshape(N, F)
        batch["sigma"] = sigma
        batch[f"D_{variant}"] = np.packbits(D[: N * F].astype(np.uint8).reshape(N, F), axis=0)
        batch[f"flag_{variant}"] = D[N * F:].astype(np.uint8)
        batch[f"iters_{variant}"] = np.int32(it)
        batch["seed_after"] = seed
        os.remove(os.path.join(ROOT, "oracle", "_ref", f"libbldpc_ref_C1_{variant}_F16.so"))
    np.savez_compressed(os.path.join(OUT, "binary_C1_batch16.npz"), **batch)
    # Simulation_GPU counters
    for variant, snrs in (("literal", (3.0, 5.0, 7.0)), ("fixed", (2.5, 3.0))):
        lib = build(f"C1_{variant}", 4, 24, 96, "J4_L24_Z96_BlockH.txt", 256, 10, 1024, 1, variant)
        for snr in snrs:
            c = np.zeros(6, np.int64)
            lib.ref_sim_point(snr, c.ctypes.data_as(C.POINTER(C.c_long)))
            meta["sim_points"].append({"code": "J4_L24_Z96_BlockH.txt", "variant": variant, "snrtype": 1,
                                       "snr_db": snr, "F": 256, "maxit": 10, "leastTestFrames": 1024,
                                       "counters": c.tolist()})
    with open(os.path.join(OUT, "binary_ref.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote golden fixtures to", OUT)


if __name__ == "__main__":
    sys.exit(main())
