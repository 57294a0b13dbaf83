/*
 * mixgan_b200_probe.h — test/diagnostic entry points of libmixgan_b200_dbg.so ONLY (built from the product sources with
 * -DMGB_DEBUG_BUILD plus csrc/umma_probe.cu).  They pin the tcgen05 shared-memory descriptor conventions the product
 * kernels rely on (tests/test_umma_probe.py) and measure issue / copy rates (scripts/umma_rate.py, scripts/bulk_rate.py).
 * Unlike the product ABI these calls may allocate and synchronise.  The product library exports none of them.
 */
#ifndef MIXGAN_B200_PROBE_H_
#define MIXGAN_B200_PROBE_H_

#include "mixgan_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * tcgen05 descriptor probe (used by tests/test_umma_probe.py to pin the shared-memory
 * descriptor conventions the bf16 path relies on).  Copies `a_bytes`/`b_bytes` raw bytes into
 * shared memory, issues `ksteps` tcgen05.mma (M=128, N=n, K=16, bf16 -> fp32) with the given
 * descriptor fields (all byte quantities, multiples of 16), and writes D[128][n] fp32.
 * use_bulk_copy: bit 0 = stage the images with cp.async.bulk, bit 1 = the A operand is MN-major (instruction
 * descriptor bit 15), bit 2 = the B operand is MN-major (bit 16).
 */
int mgb_probe_umma(const void* a_img, int a_bytes, const void* b_img, int b_bytes,
                   int a_start, int a_lbo, int a_sbo, int a_kadv,
                   int b_start, int b_lbo, int b_sbo, int b_kadv,
                   int n, int ksteps, int use_bulk_copy, float* d_out, int* status_out, void* stream);

/* Same probe for a CTA pair (cta_group::2): M = 256 (a_img holds two 128-row images back to back),
 * B rows split in halves (b_img holds two n/2-row images back to back); d_out is [256][n]. */
int mgb_probe_umma_2cta(const void* a_img, int a_bytes, const void* b_img, int b_bytes,
                        int a_lbo, int a_sbo, int a_kadv, int b_lbo, int b_sbo, int b_kadv,
                        int n, int ksteps, float* d_out, int* status_out, void* stream);

/* tcgen05 issue-rate probe (scripts/umma_rate.py): `grid` CTAs (CTA pairs when cta2 != 0) each issue
 * reps*ksteps tcgen05.mma (M = 128, or 256 for a pair; N = n; K = 16) on zeroed shared-memory operands with
 * the given K-major descriptor fields and accumulate into nacc rotating TMEM tiles; cycles_out[i] (device,
 * int64, one per launched CTA) receives the SM cycles from the first issue to the completion of the commit. */
int mgb_probe_umma_rate(int cta2, int grid, int n, int ksteps, int reps, int nacc, int a_lbo, int a_sbo,
                        int a_kadv, int b_lbo, int b_sbo, int b_kadv, int b_off, long long* cycles_out,
                        int* status_out, void* stream);

/* mgb_probe_umma_rate with a choice of operand contents (fill: 0 = zeros, 1 = pseudo-random values in (-1, 1)) and operand
 * format (fp16: 0 = bf16, 1 = fp16). */
int mgb_probe_umma_rate_data(int cta2, int grid, int n, int ksteps, int reps, int nacc, int a_lbo, int a_sbo,
                             int a_kadv, int b_lbo, int b_sbo, int b_kadv, int b_off, int fill, int fp16,
                             long long* cycles_out, int* status_out, void* stream);

/* Bulk-copy ingest-rate probe (scripts/bulk_rate.py): `grid` CTAs each stream iters * copies_per_slot copies of copy_bytes
 * from `src` (device, src_bytes long, L2-resident when small) into a `slots`-deep shared-memory ring; cycles_out[i] = SM
 * cycles CTA i needed.  Measures how the per-SM shared-memory fill rate depends on the size of one cp.async.bulk. */
int mgb_probe_bulk_rate(const void* src, long long src_bytes, int grid, int copy_bytes, int copies_per_slot,
                        int slots, int iters, long long* cycles_out, int* status_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MIXGAN_B200_PROBE_H_ */
