// ORACLE (test infrastructure only -- never linked into or called by the product path).
//
// java.util.Random restated from its published specification (JDK javadoc of java.util.Random:
// 48-bit LCG, next(bits), nextInt(n), nextFloat, nextDouble, nextGaussian). The reference seeds one
// instance per experiment (src/experiment/Experiment.java:22 `new Random(seed)`) and shares it between
// arena placement (src/arena/RoundedRectangleEnclosure.java:60-61, src/arena/Arena.java:159,178) and
// every controller (src/controllers/CacheConsController.java:206, BHDController.java:69,
// src/controllers/cluster/ExtremaSelector.java:33). SURVEY App. D.
#pragma once
#include <cstdint>
#include <cmath>

namespace oracle {

struct JavaRandom {
    uint64_t seed = 0;
    bool haveNextNextGaussian = false;
    double nextNextGaussian = 0.0;

    JavaRandom() {}
    explicit JavaRandom(int64_t s) { setSeed(s); }
    void setSeed(int64_t s) {
        seed = ((uint64_t)s ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);
        haveNextNextGaussian = false;
    }
    int32_t next(int bits) {
        seed = (seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
        return (int32_t)((int64_t)seed >> (48 - bits));   // seed < 2^48, so the shift is logical
    }
    int32_t nextInt() { return next(32); }
    int32_t nextInt(int32_t n) {
        if ((n & -n) == n) return (int32_t)(((int64_t)n * (int64_t)next(31)) >> 31);
        int32_t bits, val;
        do {
            bits = next(31);
            val = bits % n;
        } while ((int32_t)((uint32_t)bits - (uint32_t)val + (uint32_t)(n - 1)) < 0);
        return val;
    }
    float nextFloat() { return next(24) / (float)(1 << 24); }
    double nextDouble() {
        int64_t hi = (int64_t)next(26) << 27;
        return (double)(hi + next(27)) * (1.0 / (double)(1LL << 53));
    }
    // Marsaglia polar method with the cached second deviate; Java uses StrictMath.sqrt/log.
    double nextGaussian() {
        if (haveNextNextGaussian) {
            haveNextNextGaussian = false;
            return nextNextGaussian;
        }
        double v1, v2, s;
        do {
            v1 = 2 * nextDouble() - 1;
            v2 = 2 * nextDouble() - 1;
            s = v1 * v1 + v2 * v2;
        } while (s >= 1 || s == 0);
        double multiplier = std::sqrt(-2 * std::log(s) / s);
        nextNextGaussian = v2 * multiplier;
        haveNextNextGaussian = true;
        return v1 * multiplier;
    }
};

}  // namespace oracle
