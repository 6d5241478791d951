// Hand-off between the scalar QL kernel for large Fock dimensions (qtet_big.cu) and the replay mode of the
// generic kernel (qtet_generic.cu).
#pragma once
#include "qtet_common.cuh"

namespace qtet {

struct BigStream {
    // written by the scalar QL kernel, read by the replay
    double2* stream = nullptr;     // [nb][capit][32]  one 512 B row per plane of a round: (c, s) per lane
    unsigned* hdr = nullptr;       // [nb][maxr][32]   l (0xff = idle) | mtop << 8 | lbot << 16 | m << 24 (the lane's own block top)
    int* it0 = nullptr;            // [nb][maxr]       first stream row of the round
    int* nrounds = nullptr;        // [nb]
    double* ev = nullptr;          // [n][d]           eigenvalues (shifted by the mean of the diagonal)
    int capit = 0, maxr = 0;
    // dense H only: Householder tridiagonalisation (generic kernel, mode 1)
    double* tri_d = nullptr;       // [n][d]
    double* tri_e = nullptr;       // [n][d]
    double* qmat = nullptr;        // [n][d][d]
    // direct path (eigenvectors by two-sided recurrences): written by the eigenvalue-only QL kernel
    double* adiag = nullptr;       // [n][d]           diagonal of the (Householder-reduced) tridiagonal minus its mean
    // staged Householder reduction (qtet_hhfrag.cu): reflectors of the first stage, hh_frag_scratch_bytes(d) per point (or null)
    double* rfl = nullptr;
};

// register-resident kernels (qtet_bigreg.cu); they store / expect Q column-major
int bigreg_padded_dim(int d);
int bigreg_max_timesteps();
int launch_hh_reg(const ProblemView& pv, const double* chis, int64_t n, const BigStream& bs, int sms, cudaStream_t st);
// the same reduction with the matrix in 8 x 8 register tiles (qtet_hhfrag.cu); launch_hh_reg forwards to it unless QTET_HH_FRAG=0
int launch_hh_frag(const ProblemView& pv, const double* chis, int64_t n, const BigStream& bs, int sms, cudaStream_t st);
size_t hh_frag_scratch_bytes(int d);
int launch_replay_reg(const ProblemView& pv, const double* chis, int64_t n, int site, const BigStream& bs,
                      const unsigned short* pair_tab, double* loss, double* grad, int32_t* tstar, double* trace, int sms,
                      cudaStream_t st, bool direct = false, int* fb_list = nullptr, int* fb_count = nullptr);
// largest padded dimension for which the direct kernels exist (dense H: the V = Q Z tensor-core product must fit the registers)
int bigreg_direct_max_dim(bool dense);

int launch_generic_mode(const ProblemView& pv, const double* chis, int64_t B, int site, double* loss, double* grad,
                        int32_t* tstar, double* trace, cudaStream_t st, int mode, const BigStream& bs);

}  // namespace qtet
