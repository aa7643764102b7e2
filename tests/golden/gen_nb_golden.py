#!/usr/bin/env python
"""Generates tests/golden/nb_*.npz / nb_ref.json from the reference's OWN non-binary CPU decoder.

Build container only (needs /root/reference):  python tests/golden/gen_nb_golden.py
oracle/build_ref_nb.sh compiles myNBLDPC/src/{LDPC_Decoder,GF,LDPC_Encoder,Simulation,struct}.cpp
UNCHANGED; this script drives Modulate + AWGNChannel_CPU (seeds 173,173,173) + Demodulate and
Decoding_EMS / Decoding_TMM / Decoding_layered_TMM frame by frame and records: the channel
samples of every frame, L_ch of the first two frames, and per algorithm the decoder's return value,
iter_number and all decided symbols.  Coefficients are read RAW, exactly as the reference does
(for *_exp.txt files that means the exponent is used as the element) — these vectors
pin the algorithms, not the exponent mapping.
"""
import ctypes as C, hashlib, json, os, re, subprocess, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/myNBLDPC"
OUT = os.path.join(ROOT, "tests", "golden")

CONFIGS = [
    # name, matrix, constellation, n_qam, q, dc, dv, snr_db, algos, frames
    ("BDS", "BDS.576.288.GF.64.txt", "./Constellation/BPSK.txt", 2, 64, 4, 2, 2.5, (0, 1, 3), 12),
    ("C4", "LDPC_N576_K288_GF64_d1_exp.txt", "./Constellation/GRAY_64QAM.txt", 64, 64, 4, 2, 9.0, (0,), 10),
    ("C5", "LDPC_N576_K480_GF256_exp.txt", "./Constellation/BPSK.txt", 2, 256, 12, 2, 4.0, (1, 3), 6),
]


def main():
    os.chdir(REF)
    meta = {"gf_sha256": {}, "configs": {}}
    for q in (4, 8, 16, 32, 64, 128, 256, 512):
        txt = open(f"{REF}/GF/Arith.Table.GF.{q}.txt").read().split()
        nums = [int(t) for t in txt if re.fullmatch(r"-?\d+", t)]
        # header "GF(q) with Primitive Polynomial: P." -> tokens with trailing '.' are skipped; the
        # numeric body is mul (q*q), add (q*q), inverse (q)
        body = np.array(nums[-(2 * q * q + q):], np.int32)
        meta["gf_sha256"][str(q)] = {"mul": hashlib.sha256(body[: q * q].tobytes()).hexdigest(),
                                     "inv": hashlib.sha256(body[2 * q * q:].tobytes()).hexdigest()}
    cw = open(f"{REF}/include/codeword_test.h").read()
    meta["CodeWord_sym_test"] = [int(x) for x in re.search(r"\{([^}]*)\}", cw).group(1).split(",")]
    for name, matrix, cfile, nqam, q, dc, dv, snr, algos, frames in CONFIGS:
        subprocess.check_call([f"{ROOT}/oracle/build_ref_nb.sh", f"gold_{name}", matrix, cfile, str(nqam), str(q),
                               str(dc), str(dv), "2", "2", "20"])
        ref = C.CDLL(f"{ROOT}/oracle/_ref/libnbldpc_ref_gold_{name}.so")
        assert ref.nbref_init() == 0
        ref.nbref_sigma.restype = C.c_float
        ref.nbref_sigma.argtypes = [C.c_float]
        ref.nbref_channel.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        ref.nbref_decode.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        info = np.zeros(8, np.int32)
        ref.nbref_info(info.ctypes.data_as(C.POINTER(C.c_int)))
        N, M, qq, p = map(int, info[:4])
        sym = np.array(meta["CodeWord_sym_test"], np.int32) if name == "BDS" else np.zeros(N, np.int32)
        sigma = ref.nbref_sigma(snr)
        L = N * p if nqam == 2 else N
        seed = np.array([173, 173, 173], np.int32)
        rx = np.zeros((frames, 2 * L), np.float32)
        lch2 = np.zeros((2, N * (q - 1)), np.float32)
        res = {a: {"ret": [], "iter": [], "out": []} for a in algos}
        for f in range(frames):
            lch = np.zeros(N * (q - 1), np.float32)
            ref.nbref_channel(seed.ctypes.data, sigma, sym.ctypes.data, rx[f].ctypes.data, lch.ctypes.data)
            if f < 2:
                lch2[f] = lch
            for a in algos:
                o = np.zeros(N, np.int32)
                it = C.c_int(0)
                r = ref.nbref_decode(a, lch.ctypes.data, o.ctypes.data, C.byref(it))
                res[a]["ret"].append(r); res[a]["iter"].append(it.value); res[a]["out"].append(o.copy())
        arrays = {"rx": rx, "L_ch_first2": lch2, "sym": sym, "sigma": np.float32(sigma), "seed_after": seed}
        for a in algos:
            arrays[f"ret_{a}"] = np.array(res[a]["ret"], np.int32)
            arrays[f"iter_{a}"] = np.array(res[a]["iter"], np.int32)
            arrays[f"out_{a}"] = np.array(res[a]["out"], np.int16)
        np.savez_compressed(os.path.join(OUT, f"nb_{name}.npz"), **arrays)
        meta["configs"][name] = {"matrix": matrix, "constellation": cfile, "n_qam": nqam, "q": q, "dc_max": dc,
                                 "dv_max": dv, "snr_db": snr, "algos": list(algos), "frames": frames, "maxit": 20,
                                 "ems_nm": 2, "ems_nc": 2, "coef_is_exponent": 0}
        os.remove(f"{ROOT}/oracle/_ref/libnbldpc_ref_gold_{name}.so")
    with open(os.path.join(OUT, "nb_ref.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote NB golden fixtures")


if __name__ == "__main__":
    sys.exit(main())
