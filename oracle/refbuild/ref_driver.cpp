// C entry points over the REFERENCE library (Elemental itself, compiled by the
// Makefile beside this file).  TEST INFRASTRUCTURE: used by tests/ to pin the
// numpy oracle and by bench.py's cpu_baseline / --impl reference legs only.
//
// Every call wraps host column-major buffers in DistMatrix<T,MC,MR> objects on
// the reference's default Grid (1x1 under the single-process MPI shim) and runs
// the reference's own distributed code path at the requested Blocksize(), the
// same way its tests do on one rank (tests/blas_like/Gemm.cpp,
// tests/lapack_like/Cholesky.cpp).
#include <El.hpp>
#include <cstring>
#include <string>

extern "C" {
void scipy_openblas_set_num_threads(int);
int scipy_openblas_get_num_threads(void);
char* scipy_openblas_get_corename(void);
char* scipy_openblas_get_config(void);
}

namespace {
bool g_init = false;
std::string g_err;

El::Orientation orient(char c) {
    switch (c) {
        case 'N': case 'n': return El::NORMAL;
        case 'T': case 't': return El::TRANSPOSE;
        default: return El::ADJOINT;
    }
}
El::UpperOrLower ul(char c) { return (c == 'L' || c == 'l') ? El::LOWER : El::UPPER; }
El::LeftOrRight lr(char c) { return (c == 'L' || c == 'l') ? El::LEFT : El::RIGHT; }
El::UnitOrNonUnit un(char c) { return (c == 'U' || c == 'u') ? El::UNIT : El::NON_UNIT; }

template <class F>
int guarded(F&& f) {
    try {
        if (!g_init) { int argc = 0; char** argv = nullptr; El::Initialize(argc, argv); g_init = true; }
        f();
        return 0;
    } catch (const El::NonHPDMatrixException& e) { g_err = e.what(); return 2; }
    catch (const El::SingularMatrixException& e) { g_err = e.what(); return 3; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
}

template <class T>
void attach(El::DistMatrix<T>& M, El::Int m, El::Int n, T* buf, El::Int ld) {
    M.Attach(m, n, El::Grid::Default(), 0, 0, buf, ld, 0);
}
template <class T>
void lattach(El::DistMatrix<T>& M, El::Int m, El::Int n, const T* buf, El::Int ld) {
    M.LockedAttach(m, n, El::Grid::Default(), 0, 0, buf, ld, 0);
}

template <class T>
int gemm(char oA, char oB, int m, int n, int k, T alpha, const T* A, int lda, const T* B, int ldb,
         T beta, T* C, int ldc, int nb, int alg) {
    return guarded([&] {
        El::PushBlocksizeStack(nb);
        El::DistMatrix<T> dA, dB, dC;
        const bool ta = (oA != 'N' && oA != 'n'), tb = (oB != 'N' && oB != 'n');
        lattach(dA, ta ? k : m, ta ? m : k, A, lda);
        lattach(dB, tb ? n : k, tb ? k : n, B, ldb);
        attach(dC, m, n, C, ldc);
        El::Gemm(orient(oA), orient(oB), alpha, dA, dB, beta, dC, El::GemmAlgorithm(alg));
        El::PopBlocksizeStack();
    });
}
template <class T>
int cholesky(char uplo, int n, T* A, int lda, int nb) {
    return guarded([&] {
        El::PushBlocksizeStack(nb);
        El::DistMatrix<T> dA;
        attach(dA, n, n, A, lda);
        try { El::Cholesky(ul(uplo), dA); } catch (...) { El::PopBlocksizeStack(); throw; }
        El::PopBlocksizeStack();
    });
}
template <class T>
int hpdsolve(char uplo, char o, int n, int nrhs, const T* A, int lda, T* B, int ldb, int nb) {
    return guarded([&] {
        El::PushBlocksizeStack(nb);
        El::DistMatrix<T> dA, dB;
        lattach(dA, n, n, A, lda);
        attach(dB, n, nrhs, B, ldb);
        try { El::HPDSolve(ul(uplo), orient(o), dA, dB); } catch (...) { El::PopBlocksizeStack(); throw; }
        El::PopBlocksizeStack();
    });
}
template <class T>
int trsm(char side, char uplo, char o, char diag, int m, int n, T alpha, const T* A, int lda, T* B,
         int ldb, int nb) {
    return guarded([&] {
        El::PushBlocksizeStack(nb);
        El::DistMatrix<T> dA, dB;
        const int ka = (side == 'L' || side == 'l') ? m : n;
        lattach(dA, ka, ka, A, lda);
        attach(dB, m, n, B, ldb);
        El::Trsm(lr(side), ul(uplo), orient(o), un(diag), alpha, dA, dB);
        El::PopBlocksizeStack();
    });
}
template <class T>
int herk(char uplo, char o, int n, int k, El::Base<T> alpha, const T* A, int lda, El::Base<T> beta,
         T* C, int ldc, int nb) {
    return guarded([&] {
        El::PushBlocksizeStack(nb);
        El::DistMatrix<T> dA, dC;
        const bool tr = (o != 'N' && o != 'n');
        lattach(dA, tr ? k : n, tr ? n : k, A, lda);
        attach(dC, n, n, C, ldc);
        El::Herk(ul(uplo), orient(o), alpha, dA, beta, dC);
        El::PopBlocksizeStack();
    });
}
template <class T>
int trrk(char uplo, char oA, char oB, int n, int k, T alpha, const T* A, int lda, const T* B,
         int ldb, T beta, T* C, int ldc, int nb) {
    return guarded([&] {
        El::PushBlocksizeStack(nb);
        El::DistMatrix<T> dA, dB, dC;
        const bool ta = (oA != 'N' && oA != 'n'), tb = (oB != 'N' && oB != 'n');
        lattach(dA, ta ? k : n, ta ? n : k, A, lda);
        lattach(dB, tb ? n : k, tb ? k : n, B, ldb);
        attach(dC, n, n, C, ldc);
        El::Trrk(ul(uplo), orient(oA), orient(oB), alpha, dA, dB, beta, dC);
        El::PopBlocksizeStack();
    });
}
// ---- the level-3 siblings SURVEY.md section 8f ranks next (they funnel into the same leaves) ----
template <class T>
int syr2k(char uplo, char o, int n, int k, T alpha, const T* A, int lda, const T* B, int ldb, T beta, T* C,
          int ldc, int conjugate, int nb) {
    return guarded([&] {
        El::PushBlocksizeStack(nb);
        El::DistMatrix<T> dA, dB, dC;
        const bool tr = (o != 'N' && o != 'n');
        lattach(dA, tr ? k : n, tr ? n : k, A, lda);
        lattach(dB, tr ? k : n, tr ? n : k, B, ldb);
        attach(dC, n, n, C, ldc);
        El::Syr2k(ul(uplo), orient(o), alpha, dA, dB, beta, dC, conjugate != 0);
        El::PopBlocksizeStack();
    });
}
template <class T>
int symm(char side, char uplo, int m, int n, T alpha, const T* A, int lda, const T* B, int ldb, T beta, T* C,
         int ldc, int conjugate, int nb) {
    return guarded([&] {
        El::PushBlocksizeStack(nb);
        El::DistMatrix<T> dA, dB, dC;
        const int ka = (side == 'L' || side == 'l') ? m : n;
        lattach(dA, ka, ka, A, lda);
        lattach(dB, m, n, B, ldb);
        attach(dC, m, n, C, ldc);
        El::Symm(lr(side), ul(uplo), alpha, dA, dB, beta, dC, conjugate != 0);
        El::PopBlocksizeStack();
    });
}
template <class T>
int trmm(char side, char uplo, char o, char diag, int m, int n, T alpha, const T* A, int lda, T* B, int ldb,
         int nb) {
    return guarded([&] {
        El::PushBlocksizeStack(nb);
        El::DistMatrix<T> dA, dB;
        const int ka = (side == 'L' || side == 'l') ? m : n;
        lattach(dA, ka, ka, A, lda);
        attach(dB, m, n, B, ldb);
        El::Trmm(lr(side), ul(uplo), orient(o), un(diag), alpha, dA, dB);
        El::PopBlocksizeStack();
    });
}
// ---- section 8f rank 3: LU (src/lapack_like/factor/LU.cpp), pivoted Cholesky, CholeskyMod, LinearSolve ----
struct BlockScope {
    explicit BlockScope(int nb) { El::PushBlocksizeStack(nb); }
    ~BlockScope() { El::PopBlocksizeStack(); }
};
// preimage[i] = the row of the input that ends up as row i of P A (DistPermutation::Preimage)
void preimages(const El::DistPermutation& P, int n, int* out) {
    for (int i = 0; i < n; ++i) out[i] = (int)P.Preimage(i);
}
template <class T>
int lu_nopiv(int m, int n, T* A, int lda, int nb) {
    return guarded([&] {
        BlockScope b(nb);
        El::DistMatrix<T> dA;
        attach(dA, m, n, A, lda);
        El::LU(dA);
    });
}
template <class T>
int lu_piv(int m, int n, T* A, int lda, int* preimage, int nb) {
    return guarded([&] {
        BlockScope b(nb);
        El::DistMatrix<T> dA;
        attach(dA, m, n, A, lda);
        El::DistPermutation P(El::Grid::Default());
        El::LU(dA, P);
        preimages(P, m, preimage);
    });
}
// factor a copy of A with partial pivoting, then lu::SolveAfter(orientation, LU, P, B)
template <class T>
int lu_piv_solve(char o, int n, int nrhs, const T* A, int lda, T* B, int ldb, int nb) {
    return guarded([&] {
        BlockScope b(nb);
        El::DistMatrix<T> dA, dB, F;
        lattach(dA, n, n, A, lda);
        attach(dB, n, nrhs, B, ldb);
        F = dA;
        El::DistPermutation P(El::Grid::Default());
        El::LU(F, P);
        El::lu::SolveAfter(orient(o), F, P, dB);
    });
}
template <class T>
int linear_solve(int n, int nrhs, const T* A, int lda, T* B, int ldb, int nb) {
    return guarded([&] {
        BlockScope b(nb);
        El::DistMatrix<T> dA, dB;
        lattach(dA, n, n, A, lda);
        attach(dB, n, nrhs, B, ldb);
        El::LinearSolve(dA, dB);
    });
}
template <class T>
int cholesky_piv(char uplo, int n, T* A, int lda, int* preimage, int nb) {
    return guarded([&] {
        BlockScope b(nb);
        El::DistMatrix<T> dA;
        attach(dA, n, n, A, lda);
        El::DistPermutation P(El::Grid::Default());
        El::Cholesky(ul(uplo), dA, P);
        preimages(P, n, preimage);
    });
}
template <class T>
int cholesky_piv_solve(char uplo, char o, int n, int nrhs, const T* A, int lda, T* B, int ldb, int nb) {
    return guarded([&] {
        BlockScope b(nb);
        El::DistMatrix<T> dA, dB, F;
        lattach(dA, n, n, A, lda);
        attach(dB, n, nrhs, B, ldb);
        F = dA;
        El::DistPermutation P(El::Grid::Default());
        El::Cholesky(ul(uplo), F, P);
        El::cholesky::SolveAfter(ul(uplo), orient(o), F, P, dB);
    });
}
// T (n x n factor, the uplo triangle) and V (n x k, overwritten) as in CholeskyMod(uplo, T, alpha, V)
template <class T>
int cholesky_mod(char uplo, int n, int k, T* Tf, int ldt, El::Base<T> alpha, T* V, int ldv, int nb) {
    return guarded([&] {
        BlockScope b(nb);
        El::DistMatrix<T> dT, dV;
        attach(dT, n, n, Tf, ldt);
        attach(dV, n, k, V, ldv);
        El::CholeskyMod(ul(uplo), dT, alpha, dV);
    });
}
typedef El::Complex<double> Z;
typedef El::Complex<float> Cf;
}  // namespace

extern "C" {
const char* elref_last_error() { return g_err.c_str(); }
void elref_set_threads(int n) { scipy_openblas_set_num_threads(n); }
int elref_get_threads() { return scipy_openblas_get_num_threads(); }
const char* elref_blas_corename() { return scipy_openblas_get_corename(); }
const char* elref_blas_config() { return scipy_openblas_get_config(); }
int elref_blocksize() { return guarded([] {}) == 0 ? (int)El::Blocksize() : -1; }

int elref_gemm_d(char a, char b, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C, int ldc, int nb, int alg) { return gemm<double>(a, b, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, nb, alg); }
int elref_gemm_s(char a, char b, int m, int n, int k, float alpha, const float* A, int lda, const float* B, int ldb, float beta, float* C, int ldc, int nb, int alg) { return gemm<float>(a, b, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, nb, alg); }
int elref_gemm_z(char a, char b, int m, int n, int k, const double* alpha, const void* A, int lda, const void* B, int ldb, const double* beta, void* C, int ldc, int nb, int alg) { return gemm<Z>(a, b, m, n, k, Z(alpha[0], alpha[1]), (const Z*)A, lda, (const Z*)B, ldb, Z(beta[0], beta[1]), (Z*)C, ldc, nb, alg); }
int elref_gemm_c(char a, char b, int m, int n, int k, const float* alpha, const void* A, int lda, const void* B, int ldb, const float* beta, void* C, int ldc, int nb, int alg) { return gemm<Cf>(a, b, m, n, k, Cf(alpha[0], alpha[1]), (const Cf*)A, lda, (const Cf*)B, ldb, Cf(beta[0], beta[1]), (Cf*)C, ldc, nb, alg); }

int elref_cholesky_d(char uplo, int n, double* A, int lda, int nb) { return cholesky<double>(uplo, n, A, lda, nb); }
int elref_cholesky_s(char uplo, int n, float* A, int lda, int nb) { return cholesky<float>(uplo, n, A, lda, nb); }
int elref_cholesky_z(char uplo, int n, void* A, int lda, int nb) { return cholesky<Z>(uplo, n, (Z*)A, lda, nb); }
int elref_cholesky_c(char uplo, int n, void* A, int lda, int nb) { return cholesky<Cf>(uplo, n, (Cf*)A, lda, nb); }

int elref_hpdsolve_d(char uplo, char o, int n, int nrhs, const double* A, int lda, double* B, int ldb, int nb) { return hpdsolve<double>(uplo, o, n, nrhs, A, lda, B, ldb, nb); }
int elref_hpdsolve_z(char uplo, char o, int n, int nrhs, const void* A, int lda, void* B, int ldb, int nb) { return hpdsolve<Z>(uplo, o, n, nrhs, (const Z*)A, lda, (Z*)B, ldb, nb); }

int elref_trsm_d(char side, char uplo, char o, char diag, int m, int n, double alpha, const double* A, int lda, double* B, int ldb, int nb) { return trsm<double>(side, uplo, o, diag, m, n, alpha, A, lda, B, ldb, nb); }
int elref_trsm_z(char side, char uplo, char o, char diag, int m, int n, const double* alpha, const void* A, int lda, void* B, int ldb, int nb) { return trsm<Z>(side, uplo, o, diag, m, n, Z(alpha[0], alpha[1]), (const Z*)A, lda, (Z*)B, ldb, nb); }

int elref_herk_d(char uplo, char o, int n, int k, double alpha, const double* A, int lda, double beta, double* C, int ldc, int nb) { return herk<double>(uplo, o, n, k, alpha, A, lda, beta, C, ldc, nb); }
int elref_herk_z(char uplo, char o, int n, int k, double alpha, const void* A, int lda, double beta, void* C, int ldc, int nb) { return herk<Z>(uplo, o, n, k, alpha, (const Z*)A, lda, beta, (Z*)C, ldc, nb); }

int elref_trrk_d(char uplo, char oA, char oB, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C, int ldc, int nb) { return trrk<double>(uplo, oA, oB, n, k, alpha, A, lda, B, ldb, beta, C, ldc, nb); }

int elref_syr2k_d(char uplo, char o, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C, int ldc, int conj, int nb) { return syr2k<double>(uplo, o, n, k, alpha, A, lda, B, ldb, beta, C, ldc, conj, nb); }
int elref_syr2k_z(char uplo, char o, int n, int k, const double* alpha, const void* A, int lda, const void* B, int ldb, const double* beta, void* C, int ldc, int conj, int nb) { return syr2k<Z>(uplo, o, n, k, Z(alpha[0], alpha[1]), (const Z*)A, lda, (const Z*)B, ldb, Z(beta[0], beta[1]), (Z*)C, ldc, conj, nb); }
int elref_symm_d(char side, char uplo, int m, int n, double alpha, const double* A, int lda, const double* B, int ldb, double beta, double* C, int ldc, int conj, int nb) { return symm<double>(side, uplo, m, n, alpha, A, lda, B, ldb, beta, C, ldc, conj, nb); }
int elref_symm_z(char side, char uplo, int m, int n, const double* alpha, const void* A, int lda, const void* B, int ldb, const double* beta, void* C, int ldc, int conj, int nb) { return symm<Z>(side, uplo, m, n, Z(alpha[0], alpha[1]), (const Z*)A, lda, (const Z*)B, ldb, Z(beta[0], beta[1]), (Z*)C, ldc, conj, nb); }
int elref_trmm_d(char side, char uplo, char o, char diag, int m, int n, double alpha, const double* A, int lda, double* B, int ldb, int nb) { return trmm<double>(side, uplo, o, diag, m, n, alpha, A, lda, B, ldb, nb); }
int elref_trmm_z(char side, char uplo, char o, char diag, int m, int n, const double* alpha, const void* A, int lda, void* B, int ldb, int nb) { return trmm<Z>(side, uplo, o, diag, m, n, Z(alpha[0], alpha[1]), (const Z*)A, lda, (Z*)B, ldb, nb); }
int elref_lu_d(int m, int n, double* A, int lda, int nb) { return lu_nopiv<double>(m, n, A, lda, nb); }
int elref_lu_s(int m, int n, float* A, int lda, int nb) { return lu_nopiv<float>(m, n, A, lda, nb); }
int elref_lu_z(int m, int n, void* A, int lda, int nb) { return lu_nopiv<Z>(m, n, (Z*)A, lda, nb); }
int elref_lu_c(int m, int n, void* A, int lda, int nb) { return lu_nopiv<Cf>(m, n, (Cf*)A, lda, nb); }
int elref_lu_piv_d(int m, int n, double* A, int lda, int* pre, int nb) { return lu_piv<double>(m, n, A, lda, pre, nb); }
int elref_lu_piv_s(int m, int n, float* A, int lda, int* pre, int nb) { return lu_piv<float>(m, n, A, lda, pre, nb); }
int elref_lu_piv_z(int m, int n, void* A, int lda, int* pre, int nb) { return lu_piv<Z>(m, n, (Z*)A, lda, pre, nb); }
int elref_lu_piv_c(int m, int n, void* A, int lda, int* pre, int nb) { return lu_piv<Cf>(m, n, (Cf*)A, lda, pre, nb); }
int elref_lu_piv_solve_d(char o, int n, int nrhs, const double* A, int lda, double* B, int ldb, int nb) { return lu_piv_solve<double>(o, n, nrhs, A, lda, B, ldb, nb); }
int elref_lu_piv_solve_z(char o, int n, int nrhs, const void* A, int lda, void* B, int ldb, int nb) { return lu_piv_solve<Z>(o, n, nrhs, (const Z*)A, lda, (Z*)B, ldb, nb); }
int elref_linear_solve_d(int n, int nrhs, const double* A, int lda, double* B, int ldb, int nb) { return linear_solve<double>(n, nrhs, A, lda, B, ldb, nb); }
int elref_linear_solve_z(int n, int nrhs, const void* A, int lda, void* B, int ldb, int nb) { return linear_solve<Z>(n, nrhs, (const Z*)A, lda, (Z*)B, ldb, nb); }
int elref_cholesky_piv_d(char uplo, int n, double* A, int lda, int* pre, int nb) { return cholesky_piv<double>(uplo, n, A, lda, pre, nb); }
int elref_cholesky_piv_z(char uplo, int n, void* A, int lda, int* pre, int nb) { return cholesky_piv<Z>(uplo, n, (Z*)A, lda, pre, nb); }
int elref_cholesky_piv_solve_d(char uplo, char o, int n, int nrhs, const double* A, int lda, double* B, int ldb, int nb) { return cholesky_piv_solve<double>(uplo, o, n, nrhs, A, lda, B, ldb, nb); }
int elref_cholesky_piv_solve_z(char uplo, char o, int n, int nrhs, const void* A, int lda, void* B, int ldb, int nb) { return cholesky_piv_solve<Z>(uplo, o, n, nrhs, (const Z*)A, lda, (Z*)B, ldb, nb); }
int elref_cholesky_mod_d(char uplo, int n, int k, double* T, int ldt, double alpha, double* V, int ldv, int nb) { return cholesky_mod<double>(uplo, n, k, T, ldt, alpha, V, ldv, nb); }
int elref_cholesky_mod_z(char uplo, int n, int k, void* T, int ldt, double alpha, void* V, int ldv, int nb) { return cholesky_mod<Z>(uplo, n, k, (Z*)T, ldt, alpha, (Z*)V, ldv, nb); }
}
