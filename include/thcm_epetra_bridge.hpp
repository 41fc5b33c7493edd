// =============================================================================
// thcm_epetra_bridge.hpp -- the Epetra side of the Model-API boundary.
//
// THCM::evaluate copies the Fortran CRS into its Epetra_CrsMatrix row by row (/root/reference/src/ocean/THCM.C:1052-1180:
// PutScalar(0), matrix_, per row a translation of local to global column ids + ReplaceGlobalValues with a search per entry,
// then two FillComplete calls) although the matrix was built ONCE on the static maximal graph (THCM.C:2300-2580) and only
// its values change.  JacobianBridge replaces that block: the device kernels write the values in graph order, a slot table
// built once maps every graph entry to its position in the matrix's own value storage, and fill() lands the values there --
// one device-to-host copy straight into the matrix when its storage is contiguous in this library's order (the one-rank
// case: Epetra sorts a row by local column id, which is then ascending global id = the graph order), one permutation kernel
// plus that copy otherwise.  Explicit zeros are written like everything else, so the PutScalar(0) disappears too; the
// pattern never changes, so neither FillComplete is needed again.
//
// Header-only template over the handful of Epetra_CrsMatrix / Epetra_BlockMap members it touches
//     int  NumMyRows() const;   bool Filled() const;
//     int  ExtractMyRowView(int MyRow, int& NumEntries, double*& Values, int*& Indices) const;
//     const Map& RowMap() const;  const Map& ColMap() const;      with   int Map::LID(int gid) const;  int Map::GID(int lid) const;
// so that it compiles against Trilinos unchanged (instantiate with Epetra_CrsMatrix) and, where Trilinos is absent, against
// the stand-in of tests/cpp/epetra_standin.hpp that the tests use.  Rows of the matrix that are not rows of the graph in
// length (the dense integral-condition row of SRES = 0, THCM.C:2180-2229) are left to their owner and reported.
// =============================================================================
#pragma once
#include <algorithm>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "thcm_b200.h"

namespace thcm_b200 {

template <class CrsMatrix>
class JacobianBridge {
public:
    // `A` must be fill-complete on the maximal graph; its row map must hold the owned unknowns of this rank (the standard /
    // solve map of TRIOS::Domain), its column map the owned and the ghost columns.
    JacobianBridge(thcmb_ctx* c, CrsMatrix& A) : c_(c), A_(A) {
        if (!A.Filled()) throw std::runtime_error("JacobianBridge: the matrix must be fill-complete (static maximal graph)");
        const int n = thcmb_ndim_local(c);
        nnz_ = thcmb_graph_nnz(c);
        if (nnz_ >= 2147483647LL) throw std::runtime_error("JacobianBridge: more than 2^31 - 1 graph entries on one rank");
        std::vector<int> rowptr((size_t)n + 1), col((size_t)nnz_), gid((size_t)n), hgid((size_t)std::max(thcmb_halo_size(c), 1));
        thcmb_get_graph(c, rowptr.data(), col.data());
        thcmb_local_gids(c, gid.data());
        thcmb_halo_gids(c, hgid.data());
        // value storage of the matrix: contiguous (Epetra's optimized storage: one array, rows back to back) or one array per row
        rowval_.assign((size_t)A.NumMyRows(), nullptr);
        std::vector<int> rowlen((size_t)A.NumMyRows(), 0);
        contiguous_ = true;
        for (int r = 0; r < A.NumMyRows(); r++) {
            int ne = 0; double* v = nullptr; int* idx = nullptr;
            if (A.ExtractMyRowView(r, ne, v, idx) != 0) throw std::runtime_error("JacobianBridge: ExtractMyRowView failed");
            rowval_[(size_t)r] = v; rowlen[(size_t)r] = ne;
            if (r > 0 && v != rowval_[(size_t)r - 1] + rowlen[(size_t)r - 1]) contiguous_ = false;
        }
        base_ = A.NumMyRows() > 0 ? rowval_[0] : nullptr;
        slot_.assign((size_t)nnz_, -1);
        rowj_.assign((size_t)nnz_, -1);
        mrow_.assign((size_t)n, -1);
        identity_ = contiguous_;
        std::vector<std::pair<int, int>> byg;   // (global column id, position in the matrix row)
        for (int r = 0; r < n; r++) {
            const int lr = A.RowMap().LID(gid[(size_t)r]);
            if (lr < 0) throw std::runtime_error("JacobianBridge: unknown " + std::to_string(gid[(size_t)r]) + " is not a row of the matrix on this rank");
            int ne = 0; double* v = nullptr; int* idx = nullptr;
            A.ExtractMyRowView(lr, ne, v, idx);
            const int len = rowptr[(size_t)r + 1] - rowptr[(size_t)r];
            if (ne != len) { foreign_rows_++; continue; }   // e.g. the dense integral-condition row: its owner fills it
            mrow_[(size_t)r] = lr;
            byg.resize((size_t)ne);
            for (int j = 0; j < ne; j++) byg[(size_t)j] = std::make_pair(A.ColMap().GID(idx[j]), j);
            std::sort(byg.begin(), byg.end());
            for (int q = 0; q < len; q++) {      // graph rows are sorted by global column id
                const int e = rowptr[(size_t)r] + q, lc = col[(size_t)e];
                const int g = lc < n ? gid[(size_t)lc] : hgid[(size_t)(lc - n)];
                if (byg[(size_t)q].first != g)
                    throw std::runtime_error("JacobianBridge: row " + std::to_string(gid[(size_t)r]) + " of the matrix is not the maximal-graph row");
                const int j = byg[(size_t)q].second;
                rowj_[(size_t)e] = j;
                if (contiguous_) {
                    const std::ptrdiff_t off = (v + j) - base_;
                    slot_[(size_t)e] = (int)off;
                    if (off != (std::ptrdiff_t)e) identity_ = false;
                }
            }
        }
        if (foreign_rows_ > 0) identity_ = false;   // their slots must not be overwritten by a straight copy
        if (contiguous_ && !identity_) {
            // foreign rows: park their graph entries in a scratch tail behind the matrix values so that the scatter stays a permutation
            long long tail = 0, total = 0;
            for (int r = 0; r < A.NumMyRows(); r++) total += rowlen[(size_t)r];
            for (long long e = 0; e < nnz_; e++) if (slot_[(size_t)e] < 0) slot_[(size_t)e] = (int)(total + tail++);
            stage_len_ = total + tail; copy_len_ = total;
            d_slot_ = (int*)thcmb_device_alloc(c, (long long)sizeof(int) * nnz_);
            d_stage_ = (double*)thcmb_device_alloc(c, (long long)sizeof(double) * stage_len_);
            if (!d_slot_ || !d_stage_) throw std::runtime_error(std::string("JacobianBridge: ") + thcmb_last_error());
            thcmb_h2d(c, d_slot_, slot_.data(), (long long)sizeof(int) * nnz_);
            // positions of the matrix that no graph entry maps to (foreign rows) must survive the copy: they are fetched back per row
        }
        if (!contiguous_) h_stage_.resize((size_t)nnz_);
    }
    ~JacobianBridge() {
        if (d_slot_) thcmb_device_free(c_, d_slot_);
        if (d_stage_) thcmb_device_free(c_, d_stage_);
    }
    JacobianBridge(const JacobianBridge&) = delete;
    JacobianBridge& operator=(const JacobianBridge&) = delete;

    // Jacobian at the state d_un (device pointer, owned unknowns) into the matrix: what THCM.C:1052-1180 does
    void fill(const double* d_un) {
        if (thcmb_jacobian_dev(c_, d_un) != 0) throw std::runtime_error(std::string("JacobianBridge: ") + thcmb_last_error());
        fill_from_stored();
    }
    // the values the library already holds (after THCM::evaluate(.., computeJac = true) of the mirror) into the matrix
    void fill_from_stored() {
        const double* d_val = thcmb_jacobian_values(c_);
        if (identity_) {
            thcmb_d2h(c_, base_, d_val, (long long)sizeof(double) * nnz_);
        } else if (contiguous_) {
            std::vector<std::vector<double>> keep;   // values of foreign rows (not ours to touch)
            if (foreign_rows_ > 0) save_foreign(keep);
            thcmb_scatter_values_dev(c_, nnz_, d_slot_, d_val, d_stage_);
            thcmb_d2h(c_, base_, d_stage_, (long long)sizeof(double) * copy_len_);
            if (foreign_rows_ > 0) restore_foreign(keep);
        } else {   // one array per row (Epetra before OptimizeStorage): host scatter
            thcmb_d2h(c_, h_stage_.data(), d_val, (long long)sizeof(double) * nnz_);
            thcmb_sync(c_);
            std::vector<int> rowptr((size_t)thcmb_ndim_local(c_) + 1), col((size_t)nnz_);
            thcmb_get_graph(c_, rowptr.data(), col.data());
            for (int r = 0; r < thcmb_ndim_local(c_); r++) {
                if (mrow_[(size_t)r] < 0) continue;
                double* v = rowval_[(size_t)mrow_[(size_t)r]];
                for (int e = rowptr[(size_t)r]; e < rowptr[(size_t)r + 1]; e++) v[rowj_[(size_t)e]] = h_stage_[(size_t)e];
            }
        }
        thcmb_sync(c_);
    }
    // diagonal of the mass matrix in the order of the owned unknowns (THCM.C:1160: localDiagB_[lid] = coB_[i] * mass_param)
    void mass_diagonal(double* diagB, double mass_param = 1.0) const {
        const int n = thcmb_ndim_local(c_);
        thcmb_get_cob(c_, diagB);
        if (mass_param != 1.0) for (int i = 0; i < n; i++) diagB[i] *= mass_param;
    }
    bool straight_copy() const { return identity_; }   // the matrix stores its values in this library's graph order
    bool contiguous() const { return contiguous_; }
    int foreign_rows() const { return foreign_rows_; }
    long long entries() const { return nnz_; }

private:
    void save_foreign(std::vector<std::vector<double>>& keep) {
        for (int r = 0; r < A_.NumMyRows(); r++) {
            bool ours = false;   // (few foreign rows: a linear scan over mrow_ per call would be wasteful, so mark once)
            if (owned_mark_.empty()) { owned_mark_.assign((size_t)A_.NumMyRows(), 0); for (int m : mrow_) if (m >= 0) owned_mark_[(size_t)m] = 1; }
            ours = owned_mark_[(size_t)r] != 0;
            if (ours) continue;
            int ne = 0; double* v = nullptr; int* idx = nullptr;
            A_.ExtractMyRowView(r, ne, v, idx);
            keep.emplace_back(v, v + ne);
        }
    }
    void restore_foreign(const std::vector<std::vector<double>>& keep) {
        thcmb_sync(c_);
        size_t q = 0;
        for (int r = 0; r < A_.NumMyRows(); r++) {
            if (owned_mark_[(size_t)r]) continue;
            int ne = 0; double* v = nullptr; int* idx = nullptr;
            A_.ExtractMyRowView(r, ne, v, idx);
            std::copy(keep[q].begin(), keep[q].end(), v);
            q++;
        }
    }

    thcmb_ctx* c_;
    CrsMatrix& A_;
    long long nnz_ = 0, stage_len_ = 0, copy_len_ = 0;
    bool contiguous_ = false, identity_ = false;
    int foreign_rows_ = 0;
    double* base_ = nullptr;
    std::vector<double*> rowval_;
    std::vector<int> slot_, rowj_, mrow_;
    std::vector<char> owned_mark_;
    std::vector<double> h_stage_;
    int* d_slot_ = nullptr;
    double* d_stage_ = nullptr;
};

}  // namespace thcm_b200
