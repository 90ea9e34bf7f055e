// Internal interfaces of the fp64 symmetric eigensolver (ipca.cu), shared with the large-d engine (bigd.cu).
#pragma once
#include "common.cuh"

namespace gsb {

struct Workspace {
    double *A, *dg, *e, *beta, *Vh, *lam, *Z, *lu, *xch, *evecs, *qx;
    unsigned *counter;
    unsigned char *swp;
    size_t bytes;
};
Workspace carve(void *base, int d, int c);
// top-c eigenpairs (descending) of the symmetric matrix held in w.A (destroyed); evecs rows are sign-normalised
int eig_top(const Workspace &w, int d, int c, double *evals, double *evecs, cudaStream_t st);

struct LanczosWs {
    double *QbT, *RT, *T, *C, *Linv, *WT, *H, *U, *lamH;
    void *eig_ws;
    size_t bytes;
};
LanczosWs carve_lanczos(void *base, int d, int c);
bool lanczos_applicable(int d, int c);
// top-c eigenpairs of G[d,d] from the block Krylov space of the rows of Vprev[c,d]
int eig_top_lanczos(const LanczosWs &lw, const double *G, const double *Vprev, int d, int c, double *evals, double *evecs,
                    cudaStream_t st);

}  // namespace gsb
