"""GF(2^p) arithmetic tables in the reference's file format (myNBLDPC/GF/Arith.Table.GF.<q>.txt,
parsed by GFInitial, myNBLDPC/src/GF.cpp:68-117): polynomial basis, alpha = 2, primitive polynomials
7, 11, 19, 37, 67, 137, 285, 529 for q = 4 ... 512.  The big table files are regenerated from the
polynomial instead of being copied; tests/golden/nb_ref.json pins their sha256 against the originals."""
import numpy as np

PRIMITIVE_POLY = {4: 7, 8: 11, 16: 19, 32: 37, 64: 67, 128: 137, 256: 285, 512: 529}


def gf_tables(q):
    poly = PRIMITIVE_POLY[q]
    p = q.bit_length() - 1
    mul = np.zeros((q, q), np.int32)
    for a in range(q):
        for b in range(a, q):
            r, x, y = 0, a, b
            while y:
                if y & 1:
                    r ^= x
                y >>= 1
                x <<= 1
                if x & q:
                    x ^= poly
            mul[a, b] = mul[b, a] = r
    inv = np.zeros(q, np.int32)
    for a in range(1, q):
        inv[a] = int(np.nonzero(mul[a] == 1)[0][0])
    add = np.bitwise_xor.outer(np.arange(q), np.arange(q)).astype(np.int32)
    assert p >= 1
    return mul, add, inv


def write_table_file(q, path):
    mul, add, inv = gf_tables(q)
    with open(path, "w") as f:
        f.write(f"GF({q}) with Primitive Polynomial: {PRIMITIVE_POLY[q]}. \n")
        f.write("Multiply Table:\n")
        for row in mul:
            f.write(" ".join(str(int(v)) for v in row) + " \n")
        f.write("Add Table:\n")
        for row in add:
            f.write(" ".join(str(int(v)) for v in row) + " \n")
        f.write("Inverse Table:\n")
        f.write(" ".join(str(int(v)) for v in inv) + " \n")
    return path
