// Example: Newton iterations on the THCM ocean model through the C++ mirror of the reference's interface
// (include/thcm_model.hpp over libthcm_b200.so).  Same calls as the corrector of the reference's continuation
// (src/continuation/Continuation.H:642-739): computeRHS, computeJacobian, solve, state update, norm check.
//
//   make example && ./examples/newton_step tests/golden/masks/mask_natl8
//   THCM_DATA_DIR=/path/to/i-emic/data ./examples/newton_step --xml /path/to/i-emic/test/ocean/ocean_params.xml
//       (the model built from the reference's own parameter list, include/thcm_paramlist.hpp; the land mask is looked up below
//        $THCM_DATA_DIR/mkmask)
//
// Needs a B200 (the library has no CPU fallback and says so).
#include <cstdio>
#include <cstring>
#include <vector>

#include "thcm_model.hpp"
#include "thcm_paramlist.hpp"

static int newton(thcm_b200::Ocean& ocean) {
    for (int it = 0; it < 5; it++) {
        ocean.computeRHS();
        const double fnorm = ocean.getRHS('V')->norm();
        std::printf("Newton iteration %d: ||F|| = %.6e\n", it, fnorm);
        if (fnorm < 1e-8) break;
        ocean.computeJacobian();
        auto b = ocean.getRHS('C');
        b->scale(-1.0);                                   // J dx = -F
        ocean.solve(b);
        std::printf("    GMRES: %d iterations, residual %.3e\n", ocean.lastSolve().iters, ocean.lastSolve().resid);
        ocean.getState('V')->update(1.0, *ocean.getSolution('V'), 1.0);   // x += dx
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: newton_step <mask file in the format of topo.F90:41-64> | --xml <ocean_params.xml>\n"); return 2; }
    if (argc > 2 && std::strcmp(argv[1], "--xml") == 0) {
        thcm_b200::ParameterList params = thcm_b200::parameterListFromXMLFile(argv[2]);
        auto ocean = thcm_b200::makeOcean(params);
        ocean->setPar("Combined Forcing", 0.01);
        return newton(*ocean);
    }
    const double PI = 3.14159265358979323846;
    // test/ocean/ocean_params.xml: 8 x 8 x 4 North Atlantic box
    int N = 8, M = 8, L = 4, periodic = 0, zero = 0, one = 1, iza = 2;
    double xmin = 286 * PI / 180.0, xmax = 350 * PI / 180.0, ymin = 10 * PI / 180.0, ymax = 74 * PI / 180.0, hdim = 4000.0, qz = 1.0;
    // the reference's own start-up sequence (THCM.C:328-390): m_global reads the mask, get_landm hands it back
    __m_global_MOD_initialize(&N, &M, &L, &xmin, &xmax, &ymin, &ymax, &hdim, &qz, &periodic, &zero, &zero, &one, &one, &one, &iza, &one, &one,
                              &zero, &zero, &zero, &zero, argv[1], "", "", "", "");
    std::vector<int> landm((size_t)(N + 2) * (M + 2) * (L + 2));
    __m_global_MOD_get_landm(landm.data());

    thcmb_settings s;
    thcmb_default_settings(&s);
    s.N = N; s.M = M; s.L = L; s.xmin = xmin; s.xmax = xmax; s.ymin = ymin; s.ymax = ymax; s.hdim = hdim; s.qz = qz;
    thcm_b200::SolverParameters sp;
    sp.tol = 1e-6; sp.maxit = 200; sp.restart = 200; sp.precon = 1;
    thcm_b200::Ocean ocean(s, landm.data(), sp);
    ocean.setPar("Combined Forcing", 0.01);     // a small step along the forcing branch, like the first continuation step
    ocean.setPar("Wind Forcing", 1.0);
    ocean.setPar("Temperature Forcing", 10.0);
    ocean.setPar("Salinity Forcing", 1.0);
    return newton(ocean);
}
