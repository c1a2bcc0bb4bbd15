#!/usr/bin/env python
"""Prints the Sobol direction numbers used by csrc/shade/base.h (dimensions 2-4 of Joe & Kuo's
new-joe-kuo-6 table: s/a/m = 1/0/{1}, 2/1/{1,3}, 3/1/{1,3,1}), 16 columns, MSB-aligned."""


def dirs(s, a, m_init, n=16):
    m = list(m_init)
    for i in range(s, n):
        v = m[i - s] ^ (m[i - s] << s)
        for k in range(1, s):
            if (a >> (s - 1 - k)) & 1:
                v ^= m[i - k] << k
        m.append(v)
    return [(m[k] << (31 - k)) & 0xFFFFFFFF for k in range(n)]


if __name__ == "__main__":
    for d in (dirs(1, 0, [1]), dirs(2, 1, [1, 3]), dirs(3, 1, [1, 3, 1])):
        print(", ".join("0x%08xu" % x for x in d))
