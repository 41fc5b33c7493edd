// =============================================================================
// thcm_paramlist.hpp -- the THCM constructor's reading of its parameter list (src/ocean/THCM.C:186-340, 640-795) for the C++ mirror.
// Header-only and generic over the list type: anything with Teuchos::ParameterList's accessors
//     T&  get<T>(const std::string& name, T default)      (sets the default when the entry is missing, as Teuchos does)
//     PL& sublist(const std::string& name)
//     begin() / end() over (name, entry) pairs is NOT needed: the starting parameters are looked up by THCM's own 30 names
// fits -- Teuchos::ParameterList itself, or the small stand-in of tests/cpp/test_paramlist.cpp (Teuchos is not available in this
// repository's build container).  The maintainer-side use is one line:
//     auto setup = thcm_b200::setupFromParameterList(oceanParams.sublist("THCM"), comm->MyPID(), comm->NumProc(), localRank);
//     auto thcm  = thcm_b200::makeTHCM(setup);
// Defaults = THCM::getDefaultInitParameters (THCM.C:2697-2770).  The land mask comes from the library's m_global symbols, i.e. it is
// the array the B1 boundary hands to THCM.C:389 ("Read Land Mask" + "Land Mask", else the idealised "Topography" case).
// =============================================================================
#pragma once
#include <cmath>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "thcm_b200.h"
#include "thcm_model.hpp"

namespace thcm_b200 {

struct THCMSetup {
    thcmb_settings settings;
    std::vector<int> landm;                 // GLOBAL mask (l+2)(m+2)(n+2), i fastest (m_global::get_landm)
    std::vector<double> spert;              // n*m, empty unless "Read Salinity Perturbation Mask"
    bool integralCondition = false;         // "Restoring Salinity Profile" == 0 (THCM.C:653-697)
    int Nic = -1, Mic = -1, intSign = -1;
    bool fixPressurePoints = false;
    std::string scaling = "THCM";
    std::vector<std::pair<std::string, double>> startingParameters;   // the entries that are not NaN (THCM.C:781-792)
};

// the 30 XML names of THCM::int2par (THCM.C:1891-1942), index = par.F90 index
inline const char* const* thcmParameterNames() {
    static const char* const names[31] = {
        "Time", "AL_T", "Rayleigh-Number", "Vertical Ekman-Number", "Horizontal Ekman-Number", "Rossby-Number", "MIXP", "RESC", "SPL1",
        "Salinity Homotopy", "Solar Forcing", "Horizontal Peclet-Number", "Vertical Peclet-Number", "P_VC", "LAMB", "Salinity Forcing",
        "Wind Forcing", "Temperature Forcing", "Nonlinear Factor", "Combined Forcing", "ARCL", "NLES", "IFRICB", "CONT", "Energy", "ALPC",
        "CMPR", "Flux Perturbation", "Salinity Perturbation", "MKAP", "SPL2"};
    return names;
}

template <class ParameterList>
THCMSetup setupFromParameterList(ParameterList& p, int rank = 0, int nranks = 1, int device = 0, int balance = 0) {
    const double PI_ = 3.14159265358979323846;   // src/trios/THCMdefs.H:17
    THCMSetup su;
    thcmb_settings& s = su.settings;
    thcmb_default_settings(&s);
    s.N = p.template get<int>("Global Grid-Size n", 16);
    s.M = p.template get<int>("Global Grid-Size m", 16);
    s.L = p.template get<int>("Global Grid-Size l", 16);
    const double xmin = p.template get<double>("Global Bound xmin", 286.0), xmax = p.template get<double>("Global Bound xmax", 350.0);
    const double ymin = p.template get<double>("Global Bound ymin", 10.0), ymax = p.template get<double>("Global Bound ymax", 74.0);
    if (xmin < -360.0 || xmin > 360.0 || xmax < -360.0 || xmax > 360.0 || ymin < -90.0 || ymin > 90.0 || ymax < -90.0 || ymax > 90.0)
        throw std::invalid_argument("THCM: a domain bound is outside its validator's range (THCM.C:2714-2724)");
    s.xmin = xmin * PI_ / 180.0; s.xmax = xmax * PI_ / 180.0; s.ymin = ymin * PI_ / 180.0; s.ymax = ymax * PI_ / 180.0;   // THCM.C:203-206
    s.periodic = p.template get<bool>("Periodic", false) ? 1 : 0;
    s.hdim = p.template get<double>("Depth hdim", 4000.0);
    s.qz = p.template get<double>("Grid Stretching qz", 1.0);
    int itopo = p.template get<int>("Topography", 1);
    int flat = p.template get<bool>("Flat Bottom", false) ? 1 : 0;
    int rd_mask = p.template get<bool>("Read Land Mask", false) ? 1 : 0;
    const std::string maskFile = p.template get<std::string>("Land Mask", "no_mask_specified");
    s.ih = p.template get<int>("Inhomogeneous Mixing", 0);
    s.vmix = p.template get<int>("Mixing", 1);
    s.rho_mixing = p.template get<bool>("Rho Mixing", true) ? 1 : 0;
    s.tap = p.template get<int>("Taper", 1);
    s.alphaT = p.template get<double>("Linear EOS: alpha T", 1.0e-4);
    s.alphaS = p.template get<double>("Linear EOS: alpha S", 7.6e-4);
    s.TRES = p.template get<int>("Restoring Temperature Profile", 1);
    s.SRES = p.template get<int>("Restoring Salinity Profile", 1);
    su.intSign = p.template get<int>("Salinity Integral Sign", -1);
    s.ite = p.template get<int>("Levitus T", 1);
    s.its = p.template get<int>("Levitus S", 1);
    if (p.template get<bool>("Levitus Internal T/S", false))
        throw std::invalid_argument("THCM: \"Levitus Internal T/S\" reads data files that do not ship with the reference; use set_internal_forcing");
    s.coupled_T = p.template get<int>("Coupled Temperature", 0);
    s.coupled_S = p.template get<int>("Coupled Salinity", 0);
    su.fixPressurePoints = p.template get<bool>("Fix Pressure Points", false);
    s.coriolis_on = p.template get<int>("Coriolis Force", 1);
    s.forcing_type = p.template get<int>("Forcing Type", 0);
    if (s.coupled_S == 1 && s.SRES == 1) s.SRES = 0;                                               // THCM.C:253-259
    int rd_spertm = p.template get<bool>("Read Salinity Perturbation Mask", false) ? 1 : 0;
    const std::string spertFile = p.template get<std::string>("Salinity Perturbation Mask", "no_mask_specified");
    if (std::abs(su.intSign) != 1) throw std::invalid_argument("Invalid integral sign!");          // THCM.C:265-268
    s.iza = p.template get<int>("Wind Forcing Type", 2);
    // the data-file options of m_global::get_windfield / get_temforcing / get_salforcing (global.F90:425-560): not in the reference tree
    if (s.iza < 2 || (s.coupled_T == 0 && s.ite == 0 && s.TRES != 0) || (s.coupled_S == 0 && s.its == 0 && s.SRES != 0))
        throw std::invalid_argument("THCM: \"Wind Forcing Type\" < 2 / \"Levitus T\" = 0 / \"Levitus S\" = 0 read Trenberth / Levitus data files that do "
                                    "not ship with the reference; insert the fields instead");
    s.rank = rank; s.nranks = nranks; s.device = device; s.balance = balance;
    su.Nic = p.template get<int>("Integral row coordinate i", -1);
    su.Mic = p.template get<int>("Integral row coordinate j", -1);
    su.scaling = p.template get<std::string>("Scaling", "THCM");

    // m_global::initialize + get_landm (+ get_spert): THCM.C:325-400
    int N = s.N, M = s.M, L = s.L;
    __m_global_MOD_initialize(&N, &M, &L, &s.xmin, &s.xmax, &s.ymin, &s.ymax, &s.hdim, &s.qz, &s.periodic, &itopo, &flat, &rd_mask, &s.TRES,
                              &s.SRES, &s.iza, &s.ite, &s.its, &rd_spertm, &s.coupled_T, &s.coupled_S, &s.forcing_type, maskFile.c_str(),
                              spertFile.c_str(), "", "", "");
    su.landm.resize((size_t)(N + 2) * (M + 2) * (L + 2));
    __m_global_MOD_get_landm(su.landm.data());
    if (rd_spertm) { su.spert.resize((size_t)N * M); __m_global_MOD_get_spert(su.spert.data()); }

    if (s.SRES == 0) {                                                                             // THCM.C:653-697
        if (su.Nic == -1) su.Nic = N - 1;
        if (su.Mic == -1) su.Mic = M - 1;
        const size_t midx = (size_t)(su.Nic + 1) + (size_t)(N + 2) * ((size_t)(su.Mic + 1) + (size_t)(M + 2) * (size_t)L);
        if (su.landm[midx] != 0)
            throw std::invalid_argument("Integral row coordinates (" + std::to_string(su.Nic) + "," + std::to_string(su.Mic) +
                                        ") give a land point! Please give better coordinates in xml.");
        su.integralCondition = true;
    }
    auto& start = p.sublist("Starting Parameters");
    const char* const* names = thcmParameterNames();
    for (int i = 1; i <= 30; i++) {
        const double v = start.template get<double>(names[i], std::numeric_limits<double>::quiet_NaN());
        if (!std::isnan(v)) su.startingParameters.emplace_back(names[i], v);
    }
    return su;
}

// the device part of the constructor: creates the library context (needs a CUDA device) and applies the options above
inline std::shared_ptr<THCM> makeTHCM(const THCMSetup& su) {
    auto t = std::make_shared<THCM>(su.settings, su.landm.data());
    if (!su.spert.empty()) thcmb_insert_field(t->context(), 4 /* SF_SPERT */, su.spert.data());
    if (su.integralCondition) thcmb_enable_intcond(t->context(), su.Nic, su.Mic, su.intSign);
    if (su.fixPressurePoints) thcmb_fix_pressure_points(t->context(), 1);
    for (const auto& kv : su.startingParameters) t->setParameter(kv.first, kv.second);
    return t;
}

}  // namespace thcm_b200
