/*
 * petal_oracle.c -- CPU oracle for the hot path of petabi/petal-neighbors v0.18.0.
 *
 * TEST INFRASTRUCTURE ONLY (see petal_oracle_impl.h).  A plain-C restatement of
 *   src/distance.rs:26-35            Euclidean::distance (sequential, non-FMA fold + sqrt)
 *   src/ball_tree.rs:38-63,445-613   BallTree build (complete binary tree, 1-2 point leaves)
 *   src/ball_tree.rs:80-294          query_nearest / query / query_radius traversals
 *   src/vantage_point_tree.rs:51-197 VantagePointTree build + query_nearest
 * plus a brute-force (distance, index) oracle (src/ball_tree.rs:873-894).
 *
 * Parity pinning: the reference is Rust and no Rust toolchain exists in this image, so the
 * reference itself cannot be run here.  The oracle is pinned against every known-answer vector
 * in the reference's own tests and doctests (tests/test_oracle_kat.py, SURVEY.md 8c items
 * 1-11).  Unpinned by the reference: f32 results, d > 3, index order among equal distances.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp; no -ffast-math, no -march).
 */
#include <math.h>
#include <float.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_free(void *p) { free(p); }
int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#define REAL float
#define SFX f32
#define SQRT sqrtf
#define REAL_MAX FLT_MAX
#include "petal_oracle_impl.h"
#undef REAL
#undef SFX
#undef SQRT
#undef REAL_MAX

#define REAL double
#define SFX f64
#define SQRT sqrt
#define REAL_MAX DBL_MAX
#include "petal_oracle_impl.h"
#undef REAL
#undef SFX
#undef SQRT
#undef REAL_MAX
