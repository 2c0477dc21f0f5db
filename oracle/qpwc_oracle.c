/* qpwc_oracle.c -- CPU ORACLE for the qpwcnet cost-volume / warp hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product (qpwcnet_b200) never
 * imports, links or calls anything under oracle/ and has no CPU path of its own.
 *
 * What it is: a plain-C restatement (fp32 and fp64) of the reference algorithm
 *   - CostVolume / CostVolumeV2   qpwcnet/core/layers.py:72-100, 117-132
 *   - Warp (tf_warp)              qpwcnet/core/warp.py:8-47, 63-153
 *   - WarpV2 (tfa dense_image_warp) qpwcnet/core/layers.py:171-186 (+ warp.py:156-211)
 *   - UpFlow's warp -> cost volume  qpwcnet/core/non_layers.py:377-380
 * and of the gradients TF autodiff derives for them.
 *
 * PINNING STATUS:
 *   The reference is pure Python on TensorFlow 2 + tensorflow_addons; neither is installed nor
 *   installable here, so the reference cannot be executed and `oracle/_ref` cannot be built.  The
 *   reference ships NO golden vectors, seeds or tolerances.
 *   - cost volume: the in-repo composition (layers.py:72-100) fully specifies the arithmetic and
 *     the reference's own tests pin the tfa op to it ("0.0", app/test/test_cvol_equal.py:25)
 *     => pinned by source + equivalence, not by vectors.
 *   - tf_warp: fully specified by in-repo source (warp.py:63-153).
 *   - WarpV2 / tfa interpolate_bilinear: third-party (tensorflow-addons, unpinned in setup.py:19);
 *     restated from its published algorithm.  PARITY UNPINNED except for the flow sign/axis
 *     convention (3x3 one-hot case, app/optical_flow/test_warp.py:28-33).
 *   The restatement itself is cross-checked against an independent op-by-op PyTorch-CPU
 *   transcription of the same reference lines (oracle/ref_torch.py, autograd gradients) in
 *   tests/test_oracle.py, and against hand-derived known-answer cases.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define T float
#define FN(x) x##_f32
#define FLOOR(v) floorf(v)
#include "qpwc_oracle_body.inc"
#undef T
#undef FN
#undef FLOOR

#define T double
#define FN(x) x##_f64
#define FLOOR(v) floor(v)
#include "qpwc_oracle_body.inc"
#undef T
#undef FN
#undef FLOOR

#ifdef _OPENMP
#include <omp.h>
int qo_num_threads(void) { return omp_get_max_threads(); }
void qo_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
#else
int qo_num_threads(void) { return 1; }
void qo_set_num_threads(int n) { (void)n; }
#endif
