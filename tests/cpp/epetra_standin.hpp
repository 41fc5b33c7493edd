// Stand-in for the few Epetra classes include/thcm_epetra_bridge.hpp touches (TEST INFRASTRUCTURE: Trilinos is not available in this
// environment).  Semantics follow the Epetra documentation of the members used:
//   Epetra_Map        GID(lid) / LID(gid) (-1 when absent) / NumMyElements()
//   Epetra_CrsMatrix  after FillComplete: a row's entries are sorted by LOCAL column index; OptimizeStorage() packs all values into one
//                     array (rows back to back), before it every row owns its array; ExtractMyRowView hands out pointers INTO the
//                     matrix; ReplaceGlobalValues(row gid, n, values, column gids) returns 0, or 2 when a column is not in the row
//                     ("value excluded", what THCM.C:1099-1104 tests for); PutScalar sets every stored value.
#pragma once
#include <algorithm>
#include <unordered_map>
#include <vector>

class Epetra_Map {
public:
    Epetra_Map() {}
    explicit Epetra_Map(const std::vector<int>& gids) : gid_(gids) { for (int i = 0; i < (int)gids.size(); i++) lid_[gids[i]] = i; }
    int NumMyElements() const { return (int)gid_.size(); }
    int GID(int lid) const { return (lid >= 0 && lid < (int)gid_.size()) ? gid_[lid] : -1; }
    int LID(int gid) const { auto it = lid_.find(gid); return it == lid_.end() ? -1 : it->second; }
private:
    std::vector<int> gid_;
    std::unordered_map<int, int> lid_;
};

class Epetra_CrsMatrix {
public:
    // rows: for every local row the GLOBAL column ids of its pattern (any order); the constructor plays FillComplete
    Epetra_CrsMatrix(const Epetra_Map& rowmap, const Epetra_Map& colmap, const std::vector<std::vector<int>>& rows, bool optimize_storage)
        : rowmap_(rowmap), colmap_(colmap), optimized_(optimize_storage) {
        const int n = rowmap.NumMyElements();
        ptr_.assign(n + 1, 0);
        for (int r = 0; r < n; r++) {
            std::vector<int> l;
            for (int g : rows[r]) l.push_back(colmap.LID(g));
            std::sort(l.begin(), l.end());
            ptr_[r + 1] = ptr_[r] + (int)l.size();
            idx_.insert(idx_.end(), l.begin(), l.end());
        }
        if (optimized_) all_.assign(idx_.size(), 0.0);
        else { per_row_.resize(n); for (int r = 0; r < n; r++) per_row_[r].assign(ptr_[r + 1] - ptr_[r], 0.0); }
    }
    bool Filled() const { return true; }
    bool StorageOptimized() const { return optimized_; }
    int NumMyRows() const { return rowmap_.NumMyElements(); }
    const Epetra_Map& RowMap() const { return rowmap_; }
    const Epetra_Map& ColMap() const { return colmap_; }
    int ExtractMyRowView(int r, int& n, double*& values, int*& indices) const {
        if (r < 0 || r >= NumMyRows()) return -1;
        n = ptr_[r + 1] - ptr_[r];
        values = const_cast<double*>(optimized_ ? all_.data() + ptr_[r] : per_row_[r].data());
        indices = const_cast<int*>(idx_.data() + ptr_[r]);
        return 0;
    }
    int PutScalar(double s) {
        std::fill(all_.begin(), all_.end(), s);
        for (auto& v : per_row_) std::fill(v.begin(), v.end(), s);
        return 0;
    }
    int ReplaceGlobalValues(int grow, int n, const double* values, const int* gcols) {
        const int r = rowmap_.LID(grow);
        if (r < 0) return -1;
        int ne = 0; double* v = nullptr; int* idx = nullptr;
        ExtractMyRowView(r, ne, v, idx);
        int ierr = 0;
        for (int q = 0; q < n; q++) {
            const int lc = colmap_.LID(gcols[q]);
            const int* pos = std::lower_bound(idx, idx + ne, lc);
            if (lc < 0 || pos == idx + ne || *pos != lc) { ierr = 2; continue; }
            v[pos - idx] = values[q];
        }
        return ierr;
    }
    // all stored values in (row, local column) order -- for comparisons
    std::vector<double> values() const {
        if (optimized_) return all_;
        std::vector<double> out;
        for (auto& v : per_row_) out.insert(out.end(), v.begin(), v.end());
        return out;
    }
private:
    Epetra_Map rowmap_, colmap_;
    bool optimized_;
    std::vector<int> ptr_, idx_;
    std::vector<double> all_;
    std::vector<std::vector<double>> per_row_;
};
