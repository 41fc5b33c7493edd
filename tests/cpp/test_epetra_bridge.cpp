// =============================================================================
// test_epetra_bridge.cpp -- include/thcm_epetra_bridge.hpp against the procedure it replaces.
//
// Matrix A is filled by JacobianBridge::fill (device values by slot).  Matrix B is filled the way THCM::evaluate does it
// (/root/reference/src/ocean/THCM.C:1052-1104): PutScalar(0), matrix_ through the gfortran symbol, then for every row the 1-based
// local CRS translated to global ids and handed to ReplaceGlobalValues.  Both matrices live in the Epetra stand-in
// (tests/cpp/epetra_standin.hpp) on the maximal graph; they must hold the same values bit for bit, in three storage situations:
//   1. optimized storage, Epetra's one-rank column map (identity): the bridge's straight device-to-host copy;
//   2. optimized storage, a column map in a different order (as on a rank with ghost columns): the permutation kernel;
//   3. one array per row (before OptimizeStorage): the host scatter.
// Also: a dense foreign row (the integral-condition row of SRES = 0) survives a fill untouched.
// Configuration = test/ocean/ocean_params.xml (8 x 8 x 4 North Atlantic box, mask_natl8).  Prints one JSON line.
// =============================================================================
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <vector>

#include "thcm_model.hpp"
#include "thcm_epetra_bridge.hpp"
#include "epetra_standin.hpp"

using thcm_b200::JacobianBridge;

static int N = 8, M = 8, L = 4;

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: test_epetra_bridge <mask_natl8 file>\n"); return 2; }
    const double PI = 3.14159265358979323846;
    int periodic = 0, zero = 0, one = 1, iza = 2;
    double xmin = 286 * PI / 180.0, xmax = 350 * PI / 180.0, ymin = 10 * PI / 180.0, ymax = 74 * PI / 180.0, hdim = 4000.0, qz = 1.0;
    __m_global_MOD_initialize(&N, &M, &L, &xmin, &xmax, &ymin, &ymax, &hdim, &qz, &periodic, &zero, &zero, &one, &one, &one, &iza, &one, &one,
                              &zero, &zero, &zero, &zero, argv[1], "", "", "", "");
    std::vector<int> landm((size_t)(N + 2) * (M + 2) * (L + 2));
    __m_global_MOD_get_landm(landm.data());
    // the model through the reference boundary B1 (one instance per process): init_ + set_pointers like THCM.C:619-638
    int nml = N * M * L, ih = 0, vmix = 1, tap = 1, rho_mixing = 0, coriolis_on = 1;
    double alphaT = 1.0e-4, alphaS = 7.6e-4;
    std::vector<double> z2((size_t)N * M, 0.0);
    init_(&N, &M, &L, &nml, &xmin, &xmax, &ymin, &ymax, &alphaT, &alphaS, &ih, &vmix, &tap, &rho_mixing, &coriolis_on, &periodic, landm.data(),
          z2.data(), z2.data(), z2.data(), z2.data(), z2.data());
    int nrows = 0, cap = 0;
    __m_mat_MOD_get_array_sizes(&nrows, &cap);
    std::vector<int> begA((size_t)nrows + 1), jcoA((size_t)cap), begF((size_t)nrows + 1), jcoF((size_t)nrows);
    std::vector<double> coA((size_t)cap), coB((size_t)nrows), coF((size_t)nrows);
    __m_mat_MOD_set_pointers(&nrows, &cap, begA.data(), jcoA.data(), coA.data(), coB.data(), begF.data(), jcoF.data(), coF.data());
    int idx; double val;
    idx = 19; val = 1.0; setparcs_(&idx, &val);    // Combined Forcing
    idx = 16; val = 1.0; setparcs_(&idx, &val);    // Wind
    idx = 17; val = 10.0; setparcs_(&idx, &val);   // Temperature
    idx = 15; val = 1.0; setparcs_(&idx, &val);    // Salinity
    thcmb_ctx* c = thcmb_fortran_context();
    const int n = thcmb_ndim_local(c);
    const long long nnz = thcmb_graph_nnz(c);
    std::vector<int> rowptr((size_t)n + 1), col((size_t)nnz), gid((size_t)n);
    thcmb_get_graph(c, rowptr.data(), col.data());
    thcmb_local_gids(c, gid.data());
    // state: a deterministic pattern, exact arithmetic only
    std::vector<double> x((size_t)n);
    for (int i = 0; i < n; i++) x[(size_t)i] = 0.05 * ((double)((i * 37) % 101) / 101.0 - 0.5);
    // maximal graph in global ids (one rank: local column id = position in the standard map)
    std::vector<std::vector<int>> rows((size_t)n);
    for (int r = 0; r < n; r++) for (int e = rowptr[(size_t)r]; e < rowptr[(size_t)r + 1]; e++) rows[(size_t)r].push_back(gid[(size_t)col[(size_t)e]]);
    Epetra_Map rowmap(gid);
    std::vector<int> shuffled(gid);   // a column map in another order: owned columns of the odd cells first (stands for ghost columns)
    std::stable_partition(shuffled.begin(), shuffled.end(), [](int g) { return (g / 6) % 2 == 1; });
    Epetra_Map colmap_id(gid), colmap_sh(shuffled);

    double* d_x = (double*)thcmb_device_alloc(c, (long long)sizeof(double) * n);
    thcmb_h2d(c, d_x, x.data(), (long long)sizeof(double) * n);

    long long mismatch[3] = {0, 0, 0}, excluded = 0;
    int straight[3] = {0, 0, 0};
    for (int variant = 0; variant < 3; variant++) {
        const Epetra_Map& cm = variant == 1 ? colmap_sh : colmap_id;
        Epetra_CrsMatrix A(rowmap, cm, rows, variant != 2), B(rowmap, cm, rows, variant != 2);
        // --- the bridge ---
        JacobianBridge<Epetra_CrsMatrix> bridge(c, A);
        straight[variant] = bridge.straight_copy() ? 1 : 0;
        A.PutScalar(-777.0);          // every slot must be overwritten (explicit zeros included): no PutScalar(0) needed
        bridge.fill(d_x);
        // --- the reference's procedure (THCM.C:1052-1104) ---
        B.PutScalar(0.0);
        matrix_(x.data());
        std::vector<int> indices(6 * 27 + 1);
        std::vector<double> values(6 * 27 + 1);
        for (int i = 0; i < n; i++) {
            const int index = begA[(size_t)i], numentries = begA[(size_t)i + 1] - index;
            for (int j = 0; j < numentries; j++) {
                indices[(size_t)j] = gid[(size_t)(jcoA[(size_t)(index - 1 + j)] - 1)];
                values[(size_t)j] = coA[(size_t)(index - 1 + j)];
            }
            const int ierr = B.ReplaceGlobalValues(gid[(size_t)i], numentries, values.data(), indices.data());
            if (ierr != 0) excluded++;
        }
        const std::vector<double> va = A.values(), vb = B.values();
        for (size_t q = 0; q < va.size(); q++) if (std::memcmp(&va[q], &vb[q], sizeof(double)) != 0 && !(va[q] == 0.0 && vb[q] == 0.0)) mismatch[variant]++;
    }
    // a foreign (dense) row is left alone: replace the pattern of the last S row by a dense one, as the integral condition does
    long long foreign_touched = 0; int foreign_rows = 0;
    {
        std::vector<std::vector<int>> rows2(rows);
        rows2[(size_t)n - 1].clear();
        for (int cidx = 0; cidx < n / 6; cidx++) rows2[(size_t)n - 1].push_back(6 * cidx + 5);
        Epetra_CrsMatrix A(rowmap, colmap_id, rows2, true);
        JacobianBridge<Epetra_CrsMatrix> bridge(c, A);
        foreign_rows = bridge.foreign_rows();
        A.PutScalar(-777.0);
        bridge.fill(d_x);
        int ne = 0; double* v = nullptr; int* ix = nullptr;
        A.ExtractMyRowView(n - 1, ne, v, ix);
        for (int j = 0; j < ne; j++) if (v[j] != -777.0) foreign_touched++;
        const std::vector<double> va = A.values();
        long long untouched = 0;
        for (double q : va) if (q == -777.0) untouched++;
        if (untouched != ne) foreign_touched += 1000000;   // every other slot must have been written
    }
    std::vector<double> diagB((size_t)n);
    {
        Epetra_CrsMatrix A(rowmap, colmap_id, rows, true);
        JacobianBridge<Epetra_CrsMatrix> bridge(c, A);
        bridge.mass_diagonal(diagB.data());
    }
    long long mass_mismatch = 0;
    fillcolb_();
    for (int i = 0; i < n; i++) if (diagB[(size_t)i] != coB[(size_t)i]) mass_mismatch++;
    thcmb_device_free(c, d_x);
    std::printf("{\"n\": %d, \"nnz\": %lld, \"mismatch\": [%lld, %lld, %lld], \"straight_copy\": [%d, %d, %d], \"excluded\": %lld, "
                "\"foreign_rows\": %d, \"foreign_touched\": %lld, \"mass_mismatch\": %lld}\n",
                n, nnz, mismatch[0], mismatch[1], mismatch[2], straight[0], straight[1], straight[2], excluded, foreign_rows, foreign_touched,
                mass_mismatch);
    finalize_();
    const bool ok = mismatch[0] == 0 && mismatch[1] == 0 && mismatch[2] == 0 && excluded == 0 && straight[0] == 1 && straight[1] == 0 &&
                    foreign_rows == 1 && foreign_touched == 0 && mass_mismatch == 0;
    return ok ? 0 : 1;
}
