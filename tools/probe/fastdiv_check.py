"""Brute-force check of the round-up magic division of csrc/loss_upgen.cuh (up_fastdiv): for 2 <= d <= 2^31 and n < 2^31,
n // d == (n * mul >> 32) >> shr with s = ceil(log2 d), mul = ceil(2^(31+s) / d), shr = s - 1.   python tools/probe/fastdiv_check.py"""
import random


def magic(d):
    s = (d - 1).bit_length()
    mul = ((1 << (31 + s)) + d - 1) // d
    assert mul < (1 << 32), d
    return mul, s - 1


def main():
    random.seed(1)
    ds = list(range(2, 3000)) + [random.randrange(2, 1 << 31) for _ in range(3000)]
    ds += [(1 << k) + e for k in range(1, 32) for e in (-1, 0, 1) if 2 <= (1 << k) + e <= (1 << 31)]
    bad = 0
    for d in ds:
        mul, shr = magic(d)
        ns = [0, 1, d - 1, d, d + 1, (1 << 31) - 1, (1 << 31) - 2] + [random.randrange(0, 1 << 31) for _ in range(200)]
        ns += [k * d - 1 for k in (1, 2, 3, 1000, (1 << 31) // d) if 0 < k * d - 1 < (1 << 31)]
        ns += [k * d for k in (1, 2, (1 << 31) // d - 1) if 0 < k * d < (1 << 31)]
        for n in ns:
            if n >= (1 << 31):
                continue
            bad += (((n * mul) >> 32) >> shr) != n // d
    print('divisors', len(ds), 'mismatches', bad)
    return bad


if __name__ == '__main__':
    raise SystemExit(1 if main() else 0)
