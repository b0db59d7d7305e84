/* Exhaustive check of ps::lutIndex (puckswarm_b200/csrc/ps_math.cuh): for every float x in [0, 2 pi * 1.001] the index
 * (int)(q + 0.5f) computed from q = fma(fma(-q0, p, x), 1/p, q0), q0 = x * (1/p), equals the one from the IEEE quotient
 * x / p (p = MathUtils LUT precision 0.00011f).  gcc -O2 -mfma -ffp-contract=off -o check_lut_index check_lut_index.c -lm
 * Optional argument: stride (default 1 = all 1.09e9 floats, ~25 s). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int main(int argc, char** argv) {
    const float p = 0.00011f;
    const float rp = 1.0f / p;
    uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1u, hi;
    float top = (float)(3.14159265358979323846 * 2) * 1.001f;
    memcpy(&hi, &top, 4);
    unsigned long long bad = 0, n = 0;
    for (uint64_t u = 0; u <= hi; u += stride, ++n) {
        uint32_t v = (uint32_t)u;
        float x;
        memcpy(&x, &v, 4);
        float q0 = x * rp;
        float r = fmaf(-q0, p, x);
        float q = fmaf(r, rp, q0);
        if ((int)(x / p + 0.5f) != (int)(q + 0.5f)) bad++;
    }
    printf("checked %llu floats, index mismatches %llu\n", n, bad);
    return bad != 0;
}
