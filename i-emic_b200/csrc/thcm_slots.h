// Canonical structural entries of the per-cell dependency block An(loc,row,col) (usrc.F90:609-622).
//
// The reference stores a dense 27x6x6 block per cell of which at most 104 entries can ever be
// non-zero; they are exactly the entries of the maximal graph rows (THCM.C:2320-2325: 24/22/7/11/20/20).
// Per row they are listed here sorted by (loc, col), which is the order fillcolA emits them in
// (assemble.F90:113-127: kk major, jj minor).  loc is the reference's stencil position 1..27
// (par.F90:19-26): loc = 1 + (dj+1) + 3*(di+1) + 9*lev, lev 0: k, 1: k-1, 2: k+1 (assemble.F90:157-169).
#pragma once
#include <cstdint>

namespace thcm {

struct SlotDef { int8_t loc, col; };

#define U_ 1
#define V_ 2
#define W_ 3
#define P_ 4
#define T_ 5
#define S_ 6
constexpr SlotDef SLOTS_U[24] = {{2,U_},{2,V_},{4,U_},{4,V_},{5,U_},{5,V_},{5,W_},{5,P_},{6,U_},{6,V_},{6,W_},{6,P_},
                                 {8,U_},{8,V_},{8,W_},{8,P_},{9,W_},{9,P_},{14,U_},{14,W_},{15,W_},{17,W_},{18,W_},{23,U_}};
constexpr SlotDef SLOTS_V[22] = {{2,U_},{2,V_},{4,V_},{5,U_},{5,V_},{5,W_},{5,P_},{6,V_},{6,W_},{6,P_},{8,U_},{8,V_},
                                 {8,W_},{8,P_},{9,W_},{9,P_},{14,V_},{14,W_},{15,W_},{17,W_},{18,W_},{23,V_}};
constexpr SlotDef SLOTS_W[7] = {{5,W_},{5,P_},{5,T_},{5,S_},{23,P_},{23,T_},{23,S_}};
constexpr SlotDef SLOTS_P[11] = {{1,U_},{1,V_},{2,U_},{2,V_},{4,U_},{4,V_},{5,U_},{5,V_},{5,W_},{5,P_},{14,W_}};
constexpr SlotDef SLOTS_T[20] = {{1,U_},{1,V_},{2,U_},{2,V_},{2,T_},{4,U_},{4,V_},{4,T_},{5,U_},{5,V_},{5,W_},{5,T_},
                                 {5,S_},{6,T_},{8,T_},{14,W_},{14,T_},{14,S_},{23,T_},{23,S_}};
constexpr SlotDef SLOTS_S[20] = {{1,U_},{1,V_},{2,U_},{2,V_},{2,S_},{4,U_},{4,V_},{4,S_},{5,U_},{5,V_},{5,W_},{5,T_},
                                 {5,S_},{6,S_},{8,S_},{14,W_},{14,T_},{14,S_},{23,T_},{23,S_}};
#undef U_
#undef V_
#undef W_
#undef P_
#undef T_
#undef S_

constexpr int ROW_NSLOT[6] = {24, 22, 7, 11, 20, 20};
constexpr int ROW_OFF[7] = {0, 24, 46, 53, 64, 84, 104};

template <int R> struct RowSlots;
template <> struct RowSlots<1> { static constexpr int N = 24; static constexpr const SlotDef* S = SLOTS_U; };
template <> struct RowSlots<2> { static constexpr int N = 22; static constexpr const SlotDef* S = SLOTS_V; };
template <> struct RowSlots<3> { static constexpr int N = 7;  static constexpr const SlotDef* S = SLOTS_W; };
template <> struct RowSlots<4> { static constexpr int N = 11; static constexpr const SlotDef* S = SLOTS_P; };
template <> struct RowSlots<5> { static constexpr int N = 20; static constexpr const SlotDef* S = SLOTS_T; };
template <> struct RowSlots<6> { static constexpr int N = 20; static constexpr const SlotDef* S = SLOTS_S; };

constexpr const SlotDef* row_slots(int R) {
    return R == 1 ? SLOTS_U : R == 2 ? SLOTS_V : R == 3 ? SLOTS_W : R == 4 ? SLOTS_P : R == 5 ? SLOTS_T : SLOTS_S;
}

// slot index of (loc, col) inside row R, or -1 when that entry is structurally zero
constexpr int slot_of(int R, int loc, int col) {
    const SlotDef* s = row_slots(R);
    for (int q = 0; q < ROW_NSLOT[R - 1]; q++)
        if (s[q].loc == loc && s[q].col == col) return q;
    return -1;
}

// neighbour offset of stencil position loc (assemble.F90:157-169)
constexpr int loc_di(int loc) { return ((loc - 1) % 9) / 3 - 1; }
constexpr int loc_dj(int loc) { return (loc - 1) % 3 - 1; }
constexpr int loc_dk(int loc) { return loc < 10 ? 0 : (loc < 19 ? -1 : 1); }

// position of slot q of row R inside the sorted maximal-graph row of an INTERIOR cell (nothing clipped, no periodic
// reordering): Epetra sorts a row by global column id, i.e. by (k2, j2, i2, unknown)
constexpr int interior_key(int R, int q) {
    return (((loc_dk(row_slots(R)[q].loc) + 1) * 3 + (loc_dj(row_slots(R)[q].loc) + 1)) * 3 + (loc_di(row_slots(R)[q].loc) + 1)) * 8 +
           row_slots(R)[q].col;
}
constexpr int interior_pos(int R, int q) {
    int r = 0;
    for (int p = 0; p < ROW_NSLOT[R - 1]; p++)
        if (interior_key(R, p) < interior_key(R, q)) r++;
    return r;
}

}  // namespace thcm
