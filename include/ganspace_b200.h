/*
 * ganspace_b200 -- C ABI of the B200-native activation-sampling + incremental-PCA hot path.
 *
 * The reference (harskish/ganspace) has NO C ABI / FFI on this path: its boundary is three duck-typed
 * Python surfaces.  This header is the drop-in boundary the Python host mirror
 * (ganspace_b200/*.py) binds with ctypes; each entry point cites the reference code it replaces
 * (paths relative to /root/reference).  The binding a reference maintainer would add is shown in
 * INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; `d_` pointers are CUDA device pointers, `h_` host pointers
 *   - every call is enqueued on `stream` (a cudaStream_t passed as void*), never synchronises, never
 *     allocates: the caller passes a workspace sized by the matching *_workspace_bytes() query
 *   - return 0 on success, <0 on error; gsb_last_error() returns a thread-local message
 *   - not thread-safe on one state/workspace; re-entrant across distinct ones
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns GSB_ERR_CUDA
 */
#ifndef GANSPACE_B200_H
#define GANSPACE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSB_ABI_VERSION 1

#define GSB_OK 0
#define GSB_ERR_ARG (-1)      /* bad argument (shape, alignment, null pointer) */
#define GSB_ERR_CUDA (-2)     /* CUDA runtime error (message in gsb_last_error) */
#define GSB_ERR_WORKSPACE (-3) /* workspace too small */

typedef void *gsb_stream_t; /* cudaStream_t */

int gsb_abi_version(void);
const char *gsb_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Latent sampling: NumPy legacy RandomState streams, bit-exact, one independent stream per seed.
 *   replaces  models/wrappers.py:167-175  StyleGAN2.sample_latent:
 *       rng = np.random.RandomState(seed); rng.standard_normal(512*n).reshape(n,512) -> .float()
 *   (MT19937 init_genrand seeding, 53-bit doubles, Marsaglia polar pair emitted as [f*x2, f*x1],
 *    cast to float32).  Stream s writes n_per_stream floats at d_out + s*out_stride.
 * ---------------------------------------------------------------------------------------------- */
int gsb_legacy_normal_f32(const uint32_t *d_seeds, int n_streams, int64_t n_per_stream,
                          float *d_out, int64_t out_stride, gsb_stream_t stream);

/*   replaces  models/biggan/pytorch_biggan/pytorch_pretrained_biggan/utils.py:21-33
 *       truncnorm.rvs(-2, 2, size=(B,128), random_state=RandomState(seed)).astype(f32) * truncation
 *   (SciPy draws RandomState.uniform and applies the inverse CDF: ndtri(Phi(a) + u*(Phi(b)-Phi(a)))). */
int gsb_legacy_truncnorm_f32(const uint32_t *d_seeds, int n_streams, int64_t n_per_stream,
                             double lo, double hi, float scale,
                             float *d_out, int64_t out_stride, gsb_stream_t stream);

/* raw tempered MT19937 outputs (test hook for bit-exactness of the generator itself) */
int gsb_mt19937_raw_u32(const uint32_t *d_seeds, int n_streams, int64_t n_per_stream,
                        uint32_t *d_out, int64_t out_stride, gsb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * StyleGAN2 mapping network  z[n,dim] -> w[n,dim]
 *   replaces  models/stylegan2/stylegan2-pytorch/model.py:400-409 (Generator.style =
 *             PixelNorm + n_layers x EqualLinear(dim,dim,lr_mul,activation='fused_lrelu')),
 *             model.py:14-19 (PixelNorm), :151-161 (EqualLinear.forward),
 *             op/fused_act.py:86-92 / op/fused_bias_act_kernel.cu:19-49 (bias + leaky-ReLU 0.2 * sqrt2)
 *   gsb_mapping_pack pre-multiplies weight*scale (scale = lr_mul/sqrt(dim)) and bias*lr_mul once
 *   (the reference re-materialises them on every call, model.py:153) and lays the weights out for
 *   the kernels.  d_weight: [n_layers, dim(out), dim(in)] float32, d_bias: [n_layers, dim].
 * ---------------------------------------------------------------------------------------------- */
size_t gsb_mapping_packed_bytes(int n_layers, int dim);
int gsb_mapping_pack(const float *d_weight, const float *d_bias, int n_layers, int dim, float lr_mul,
                     void *d_packed, gsb_stream_t stream);
size_t gsb_mapping_workspace_bytes(int64_t n, int dim);
/* flags: bit0 = apply PixelNorm first (always set for Generator.style); bit1 = force the SIMT fp32
 * kernels (reference-grade fp32 FMA path used to validate the tensor-core path); default = tcgen05 path
 * (fp16 hi/lo operand split, 3 MMAs per product, fp32 accumulation in TMEM) when dim % 256 == 0.
 * n_layers == 0 with bit0 set runs PixelNorm alone (d_packed may be NULL).  bits 8-15 = number of SMs the
 * persistent tensor-core kernel leaves idle for a concurrent latency-critical stream (0 = use all). */
int gsb_mapping_forward(const void *d_packed, int n_layers, int dim, const float *d_z, float *d_w,
                        int64_t n, int flags, void *d_workspace, size_t workspace_bytes,
                        gsb_stream_t stream);

/* Generic affine layer y[n,N] = x[n,K] W[N,K]^T + bias[N] in fp32 FMA (bias may be NULL: pass a workspace of
 * N floats); flags bit0 = apply sqrt2*leaky_relu_0.2.  Needs N % 128 == 0, K % 16 == 0.
 *   replaces  the nn.Linear call sites of the path outside the mapping network: BigGAN generator.gen_z
 *   (biggan model.py:211-212,232, spectral norm folded into W by the caller) and the small projections of the
 *   low-rank activation path. */
int gsb_linear_forward(const float *d_x, const float *d_w, const float *d_bias, float *d_y, int64_t n, int N,
                       int K, int flags, void *d_workspace, size_t workspace_bytes, gsb_stream_t stream);

/* Sticky status of the tensor-core path: bit0 = an activation left fp16's range (|x| > 6e4) in some call
 * since packing, i.e. the fp16 hi/lo operand split was invalid and results must be discarded (re-run with
 * flags bit1).  This is the one entry point that synchronises (device -> host copy of one word). */
int gsb_mapping_status(const void *d_packed, int n_layers, int dim, unsigned *h_flags);

/* ------------------------------------------------------------------------------------------------
 * Per-batch sufficient statistics of the incremental PCA (Gram form):
 *   mean[d] (fp64) and centred Gram (X-mean)^T (X-mean) [d,d] (fp64, full symmetric) of X[n,d] fp32
 *   (row stride ld floats).  Together with n they carry everything IncrementalPCA.partial_fit
 *   (estimators.py:68-76 -> sklearn _incremental_pca.py:254-380) extracts from the batch.
 * ---------------------------------------------------------------------------------------------- */
size_t gsb_batch_stats_workspace_bytes(int64_t n, int d);
int gsb_batch_stats(const float *d_x, int64_t n, int d, int64_t ld, double *d_mean, double *d_gram,
                    void *d_workspace, size_t workspace_bytes, gsb_stream_t stream);
/* The same for n_groups consecutive groups of rows_per_group rows (the NB-row partial_fit groups of decomposition.py:245-265)
 * in one set of launches: d_mean [n_groups][d], d_gram [n_groups][d][d].  For d % 128 == 0 (d <= 1024) both entry points
 * run on the tensor cores (tcgen05, fp16 hi/lo split operands = fp32-grade products, promoted accumulation; stats_tc.cu). */
size_t gsb_batch_stats_multi_workspace_bytes(int n_groups, int64_t rows_per_group, int d);
int gsb_batch_stats_multi(const float *d_x, int n_groups, int64_t rows_per_group, int d, int64_t ld, double *d_mean,
                          double *d_gram, void *d_workspace, size_t workspace_bytes, gsb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Incremental-PCA chain (small-d engine, d <= 1024, d % 32 == 0).
 *   replaces  estimators.py:55-81 IPCAEstimator.fit_partial/get_components, i.e. sklearn
 *   IncrementalPCA.partial_fit: running mean/var merge (extmath.py:1118-1265), rank-c truncated
 *   merge  G = V^T S^2 V + Xc^T Xc + m m^T  -> top-c eigenspace -> svd_flip sign rule; all fp64, on device.
 *   First step: symmetric eigensolve (Householder tridiagonalisation, bisection, inverse iteration,
 *   back-transform).  Later steps (d % 128 == 0, d <= 512, c % 8 == 0): the state is kept as an orthonormal
 *   basis Q of the top-c eigenspace and H = Q^T G Q (V^T S^2 V == Q H Q^T), a step is a residual-checked
 *   orthogonal iteration on one 16-CTA cluster, and gsb_ipca_export eigen-decomposes H once.  Other shapes
 *   (or GANSPACE_B200_CHAIN=direct) run the direct eigensolve every step.
 *   The state lives in device memory; chain steps must be enqueued in the reference's batch order.
 * ---------------------------------------------------------------------------------------------- */
size_t gsb_ipca_state_bytes(int d, int c);
size_t gsb_ipca_workspace_bytes(int d, int c);
int gsb_ipca_reset(void *d_state, int d, int c, gsb_stream_t stream);
/* n_seen = samples merged before this step (the host tracks it; 0 for the first step). */
int gsb_ipca_chain_step(void *d_state, int d, int c, int64_t n_seen, int64_t n_batch,
                        const double *d_mean_b, const double *d_gram_b,
                        void *d_workspace, size_t workspace_bytes, gsb_stream_t stream);
/* Export sklearn's attributes (any pointer may be NULL): components_[c,d], singular_values_[c],
 * mean_[d], var_[d], explained_variance_[c], explained_variance_ratio_[c] -- all fp64. */
int gsb_ipca_export(const void *d_state, int d, int c, int64_t n_seen,
                    double *d_components, double *d_singular_values, double *d_mean, double *d_var,
                    double *d_explained_variance, double *d_explained_variance_ratio,
                    gsb_stream_t stream);

/* Stand-alone symmetric eigensolver used by the chain (test hook): top-c eigenpairs of the fp64
 * symmetric matrix d_a[d,d] (destroyed); d_evals[c] descending, d_evecs[c,d] rows, sign-normalised.
 * d <= 1024: cluster / shared-memory tridiagonalisation; 1024 < d <= 4096: L2-resident variant. */
int gsb_sym_eig_top(double *d_a, int d, int c, double *d_evals, double *d_evecs,
                    void *d_workspace, size_t workspace_bytes, gsb_stream_t stream);

/* Sticky status word of the chain kernels (this call synchronises the stream and clears the word):
 * bit1 = a chain step (orthogonal iteration on the top-c invariant subspace) reached its
 * iteration cap before its residual tolerance -- the spectrum has no gap after component c; results since the last
 * call are not trustworthy (re-run with GANSPACE_B200_CHAIN=direct).  Checked by the host wrappers after
 * gsb_ipca_export / gsb_sym_eig_top. */
int gsb_eig_status(unsigned *h_flags, gsb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Projection statistics:  out_std[k] = population std over rows r<n of  dirs[k,:] . (x[r,:] - sub[:])
 *   replaces  decomposition.py:313-316 (random_stdevs) and :326-329 (lat_stdev).  d_sub may be NULL.
 *   Projections in fp32 FMA, moments in fp64.
 * ---------------------------------------------------------------------------------------------- */
size_t gsb_project_std_workspace_bytes(int c);
int gsb_project_std(const float *d_x, int64_t n, int d, int64_t ld, const float *d_dirs, int c,
                    const double *d_sub, float *d_out_std, void *d_workspace, size_t workspace_bytes,
                    gsb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Latent regression accumulators (Z-space runs):
 *   replaces  decomposition.py:77-139 linreg_lstsq: per batch  act = x - mean; A = (act . comp^T)/stdev;
 *   accumulates the normal equations  AtA[c,c] += A^T A,  AtZ[c,L] += A^T Z,  sumZ[L] += sum Z  (fp64)
 *   so the 1e6 x (c + L) host matrices of the reference never exist; gsb_linreg_solve returns
 *   M_t = (AtA)^-1 AtZ (Cholesky, fp64), the least-squares solution scipy.linalg.lstsq(gelsd) gives
 *   for the full-column-rank A.
 * ---------------------------------------------------------------------------------------------- */
size_t gsb_linreg_state_bytes(int c, int latent_dim);
int gsb_linreg_reset(void *d_state, int c, int latent_dim, gsb_stream_t stream);
int gsb_linreg_accumulate(void *d_state, int c, int latent_dim, const float *d_act, int64_t n, int d,
                          const float *d_comp, const float *d_mean, const float *d_stdev,
                          const float *d_z, void *d_workspace, size_t workspace_bytes,
                          gsb_stream_t stream);
size_t gsb_linreg_workspace_bytes(int64_t n, int c);
int gsb_linreg_solve(void *d_state, int c, int latent_dim, int64_t n_total, double *d_M_t,
                     double *d_z_mean, gsb_stream_t stream);
/* Rank-deficient / ill-conditioned normal equations (the reference's gelsd returns the minimum-norm solution there,
 * decomposition.py:133): gsb_linreg_solve checks its Cholesky pivots against 1e-6 of the largest diagonal entry; on failure it
 * zeroes d_M_t and records the pivot.  gsb_linreg_solve_status reads that record (0 = fine; synchronises).  The caller then
 * eigen-decomposes gsb_linreg_normal_matrix (c x c, fp64, device) with gsb_sym_eig_top and calls gsb_linreg_solve_pinv. */
int gsb_linreg_solve_status(const void *d_state, int c, int latent_dim, int *h_info, gsb_stream_t stream);
const double *gsb_linreg_normal_matrix(const void *d_state, int c, int latent_dim);
int gsb_linreg_solve_pinv(const void *d_state, int c, int latent_dim, const double *d_evals, const double *d_evecs, double rcond,
                          double *d_M_t, gsb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * StyleGAN2 synthesis up to a hooked StyledConv (layer = conv1 | convs.k; BASELINE config 5 family)
 *   replaces  models/wrappers.py:224-255 (StyleGAN2.partial_forward past 'style') over
 *             models/stylegan2/stylegan2-pytorch/model.py:181-277 (ModulatedConv2d), :280-291 (NoiseInjection),
 *             :294-304 (ConstantInput), :307-341 (StyledConv), op/fused_act.py:86-92, op/upfirdn2d.py:144-198 (Blur)
 *   for ONE global latent per sample (every layer's style is the same w, wrappers.py:202-205).
 * Layers are described in execution order: layers[0] = conv1 (on the 4x4 constant), layers[k+1] = convs.k.
 * All pointers in the descriptors are device pointers to the module's fp32 parameters in PyTorch layout.
 * gsb_synthesis_pack folds the conv scale, splits the shared weights for the tcgen05 GEMM, pre-computes
 * sum_tap w^2 for the demodulation and noise_weight*noise; gsb_synthesis_forward runs layers[0..n_run) on
 * w[n, style_dim] and writes the last layer's activation as fp32 **NHWC** rows (index (y*res + x)*cout + co;
 * the reference's NCHW flattening is a fixed permutation of it) at d_out + b*ld_out.
 * ---------------------------------------------------------------------------------------------- */
typedef struct gsb_styled_conv {
    const float *conv_weight;  /* [cout, cin, 3, 3]   ModulatedConv2d.weight[0]            model.py:222-224 */
    const float *mod_weight;   /* [cin, style_dim]    ModulatedConv2d.modulation.weight    model.py:226     */
    const float *mod_bias;     /* [cin]               ModulatedConv2d.modulation.bias (init 1)              */
    const float *act_bias;     /* [cout]              FusedLeakyReLU.bias                  fused_act.py:78  */
    const float *noise;        /* [res_out, res_out]  the fixed noise map of this layer    wrappers.py:261-267 */
    const float *noise_weight; /* [1]                 NoiseInjection.weight                model.py:284     */
    int cin, cout;             /* cin % 128 == 0, cout % 256 == 0 */
    int upsample;              /* 1: stride-2 transposed conv + blur (res_out = 2 res_in)  model.py:248-259 */
    int res_in;                /* input resolution (4 for conv1) */
} gsb_styled_conv;

size_t gsb_synthesis_packed_bytes(const gsb_styled_conv *layers, int n_layers, int style_dim);
int gsb_synthesis_pack(const gsb_styled_conv *layers, int n_layers, int style_dim, const float *d_const_input /* [cin0,4,4] */,
                       void *d_packed, size_t packed_bytes, gsb_stream_t stream);
size_t gsb_synthesis_workspace_bytes(const gsb_styled_conv *layers, int n_run, int64_t n);
int gsb_synthesis_forward(const void *d_packed, const gsb_styled_conv *layers, int n_layers, int n_run, int style_dim,
                          const float *d_w, int64_t n, float *d_out, int64_t ld_out, void *d_workspace,
                          size_t workspace_bytes, gsb_stream_t stream);
/* synchronises; *h_flags bit0 = an operand left fp16's range since the pack (results invalid) */
int gsb_synthesis_status(const void *d_packed, const gsb_styled_conv *layers, int n_layers, int style_dim, unsigned *h_flags);

/* ------------------------------------------------------------------------------------------------
 * Incremental PCA, large-d engine (conv feature maps: d up to ~10^6, where the d x d Gram of gsb_ipca_* is out of reach)
 *   replaces  estimators.py:55-81 -> sklearn IncrementalPCA.partial_fit (_incremental_pca.py:254-380) through the
 *   small side of sklearn's stacked matrix  M = [S*Vt ; X - mean_b ; mean-correction row]  (c + n_b + 1 rows of d).
 * The caller owns M: fp32 [gsb_bigd_rows(c, nb_max), d] row-major.  Rows [0, c) hold singular_values_*components_
 * (engine state); before every step the caller writes the raw batch into rows [c, c + nb) -- the producer kernels
 * write there directly -- and gsb_bigd_chain_step centres it IN PLACE, forms T = M M^T (fp32 products, fp64 sums),
 * takes its top-c eigenpairs in fp64 (direct solver; GANSPACE_B200_BIGD_CHAIN=lanczos: warm-started block Lanczos) and
 * replaces rows [0, c) by U^T M with sklearn's svd_flip signs.  d_batch_mean (optional, [d]) receives the batch mean.
 * Feature order is whatever the producer uses (PCA is equivariant under a fixed permutation of the features).
 * ---------------------------------------------------------------------------------------------- */
int gsb_bigd_rows(int c, int nb_max);
size_t gsb_bigd_state_bytes(int64_t d, int c);
#define GSB_BIGD_GRAM_TC 1   /* flags: small-side Gram on tcgen05 (fp16 hi/lo split, accumulator promoted every K = 256; needs
                                d % 64 == 0, else the fp32 FMA kernel runs); 0 = fp32 FMA kernel.  The host mirror passes
                                GSB_BIGD_GRAM_TC by default.  The same flags go to the workspace query and every phase. */
size_t gsb_bigd_workspace_bytes(int64_t d, int c, int nb_max, int flags);
int gsb_bigd_reset(void *d_state, float *d_M, int64_t d, int c, int nb_max, gsb_stream_t stream);
int gsb_bigd_chain_step(void *d_state, float *d_M, int64_t d, int c, int nb_max, int64_t n_seen, int nb, int flags,
                        double *d_batch_mean, void *d_workspace, size_t workspace_bytes, gsb_stream_t stream);
/* The same step in three phases, for feature-sharded multi-GPU runs: every rank holds a
 * column block M[:, d_r] (d = its local width) and all NB rows of the batch;
 *   gsb_bigd_step_gram    centres, merges mean/var, leaves T_r = M_r M_r^T at gsb_bigd_gram_matrix(workspace, ...)
 *                         ([rows, rows] fp64)  -> the caller all-reduces (sums) T over the ranks
 *   gsb_bigd_step_solve   eigen-solves T (redundantly on every rank), Dnew_r = U^T M_r, and returns per row
 *                         (max |.|, its signed value) over the local features in d_rowmax [c,2]
 *                         -> the caller all-gathers them and picks sklearn's svd_flip sign of the global maximum
 *   gsb_bigd_step_commit  rows [0,c) <- sign * Dnew_r (d_signs [c]; NULL = local maxima), S, sample count.
 * gsb_bigd_chain_step == gram; solve; commit(NULL). */
void *gsb_bigd_gram_matrix(void *d_workspace, int64_t d, int c, int nb_max);
int gsb_bigd_step_gram(void *d_state, float *d_M, int64_t d, int c, int nb_max, int64_t n_seen, int nb, int flags,
                       double *d_batch_mean, void *d_workspace, size_t workspace_bytes, gsb_stream_t stream);
int gsb_bigd_step_solve(void *d_state, float *d_M, int64_t d, int c, int nb_max, int64_t n_seen, int nb, int flags,
                        float *d_rowmax, void *d_workspace, size_t workspace_bytes, gsb_stream_t stream);
int gsb_bigd_step_commit(void *d_state, float *d_M, int64_t d, int c, int nb_max, int64_t n_seen, int nb, int flags,
                         const float *d_signs, void *d_workspace, size_t workspace_bytes, gsb_stream_t stream);
/* components_ [c,d] as fp32 (any pointer may be NULL); the small vectors and mean_/var_ [d] as fp64 */
int gsb_bigd_export(const void *d_state, const float *d_M, int64_t d, int c, int64_t n_seen, float *d_components,
                    double *d_singular_values, double *d_mean, double *d_var, double *d_explained_variance,
                    double *d_explained_variance_ratio, gsb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GANSPACE_B200_H */
