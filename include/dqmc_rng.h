/*
 * dqmc_rng.h -- the counter-based uniform generator that is part of the
 * dqmc_b200 ABI contract.
 *
 * The reference draws `rand()` from Julia's global RNG and only when p <= 1
 * (src/flavors/DQMC/updates/local_updates.jl:53), so its stream position is
 * data dependent and cannot be reproduced outside Julia.  The library instead
 * defines the Metropolis uniform of proposal (chain, sweep, step, site) as a
 * pure function of those indices and the context seed: Philox4x32-10 with
 *     key     = (seed_lo, seed_hi)
 *     counter = (chain, sweep, step, site)       -- chain/sweep truncated to 32 bit,
 *                                                   their high words folded into the key
 * and u = (x0 * 2^32 + x1 >> 11 ... ) mapped to [0,1) with 53 bits.
 * `step` counts the 2M slice visits of one sweep (0-based), `site` is 0-based.
 * Accept iff  p > 1  ||  u < p  -- identical decisions to the reference's
 * short-circuit form for the same u.
 *
 * Usable from C (oracle), C++ and CUDA device code.
 */
#ifndef DQMC_RNG_H
#define DQMC_RNG_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DQMC_HD __host__ __device__ __forceinline__
#else
#define DQMC_HD static inline
#endif

DQMC_HD void dqmc_philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c[0];
        const uint64_t p1 = (uint64_t)M1 * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0; k1 += W1;
    }
}

/* uniform in [0,1), 53 random bits */
DQMC_HD double dqmc_uniform(uint64_t seed, uint64_t chain, uint64_t sweep, uint32_t step,
                            uint32_t site)
{
    uint32_t c[4];
    c[0] = (uint32_t)chain; c[1] = (uint32_t)sweep; c[2] = step; c[3] = site;
    const uint32_t k0 = (uint32_t)seed ^ (uint32_t)(chain >> 32);
    const uint32_t k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(sweep >> 32);
    dqmc_philox4x32_10(c, k0, k1);
    const uint64_t bits = (((uint64_t)c[0] << 32) | (uint64_t)c[1]) >> 11;
    return (double)bits * (1.0 / 9007199254740992.0);
}

/* Second uniform of the same proposal, from the other two Philox output words: the 4-state Gauss-Hermite fields
 * (AbstractGHQField, src/flavors/DQMC/fields.jl:464-637) draw the proposed value with `rand(1:3)` before the
 * Metropolis uniform; the ABI defines that draw as dqmc_ghq_choice(x_old, dqmc_uniform_choice(...)). */
DQMC_HD double dqmc_uniform_choice(uint64_t seed, uint64_t chain, uint64_t sweep, uint32_t step,
                                   uint32_t site)
{
    uint32_t c[4];
    c[0] = (uint32_t)chain; c[1] = (uint32_t)sweep; c[2] = step; c[3] = site;
    const uint32_t k0 = (uint32_t)seed ^ (uint32_t)(chain >> 32);
    const uint32_t k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(sweep >> 32);
    dqmc_philox4x32_10(c, k0, k1);
    const uint64_t bits = (((uint64_t)c[2] << 32) | (uint64_t)c[3]) >> 11;
    return (double)bits * (1.0 / 9007199254740992.0);
}

/* x_new = choices[x_old, r] with r = 1 + floor(3 u) (fields.jl:528, 540-541, 590): the r-th of {1,2,3,4} \ {x_old}.
 * A host that owns the integer draw r passes u = (r - 0.5) / 3. */
DQMC_HD int dqmc_ghq_choice(int x_old, double u)
{
    int r = (int)(3.0 * u);
    if (r > 2) r = 2;
    if (r < 0) r = 0;
    const int c = r + 1;
    return (c >= x_old) ? c + 1 : c;
}

/* Gauss-Hermite nodes eta(x) and weights gamma(x), x = 1..4, evaluated in double arithmetic exactly like the
 * reference does (fields.jl:517-519, 579-581: s6 = sqrt(6); gammas = [1 - s6/3, 1 + s6/3, ...];
 * etas = [-sqrt(6 + 2 s6), -sqrt(6 - 2 s6), ...]); sqrt is correctly rounded on host and device. */
DQMC_HD void dqmc_ghq_tables(double eta[4], double gamma[4])
{
    const double s6 = sqrt(6.0);
    gamma[0] = 1.0 - s6 / 3.0; gamma[1] = 1.0 + s6 / 3.0; gamma[2] = 1.0 + s6 / 3.0; gamma[3] = 1.0 - s6 / 3.0;
    eta[0] = -sqrt(6.0 + 2.0 * s6); eta[1] = -sqrt(6.0 - 2.0 * s6);
    eta[2] = sqrt(6.0 - 2.0 * s6); eta[3] = sqrt(6.0 + 2.0 * s6);
}

#endif /* DQMC_RNG_H */
