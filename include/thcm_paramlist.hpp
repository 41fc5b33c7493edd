// =============================================================================
// thcm_paramlist.hpp -- the THCM constructor's reading of its parameter list (src/ocean/THCM.C:186-340, 640-795) for the C++ mirror.
// Header-only and generic over the list type: anything with Teuchos::ParameterList's accessors
//     T&  get<T>(const std::string& name, T default)      (sets the default when the entry is missing, as Teuchos does)
//     PL& sublist(const std::string& name)
//     begin() / end() over (name, entry) pairs is NOT needed: the starting parameters are looked up by THCM's own 30 names
// fits -- Teuchos::ParameterList itself, or thcm_b200::ParameterList below, which also reads the reference's XML files (Teuchos is not
// available in this repository's build container).  The maintainer-side use is one line:
//     auto setup = thcm_b200::setupFromParameterList(oceanParams.sublist("THCM"), comm->MyPID(), comm->NumProc(), localRank);
//     auto thcm  = thcm_b200::makeTHCM(setup);
// Defaults = THCM::getDefaultInitParameters (THCM.C:2697-2770).  The land mask comes from the library's m_global symbols, i.e. it is
// the array the B1 boundary hands to THCM.C:389 ("Read Land Mask" + "Land Mask", else the idealised "Topography" case).
// =============================================================================
#pragma once
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "thcm_b200.h"
#include "thcm_model.hpp"

namespace thcm_b200 {

// ---------------------------------------------------------------------------------------------------------------------
// A small typed parameter list with Teuchos::ParameterList's accessor signatures, and a reader of the reference's XML dialect
// (<ParameterList name=..> / <Parameter name=.. type=.. value=../>, nested) -- for users of the C++ mirror without Trilinos (the
// examples and tests of this repository); with Trilinos, hand setupFromParameterList a Teuchos::ParameterList instead.
// ---------------------------------------------------------------------------------------------------------------------
class ParameterList {
public:
    struct Entry { enum Type { INT, DOUBLE, BOOL, STRING } type; int i; double d; bool b; std::string s; };

private:
    std::string name_;
    std::vector<std::string> order_;                 // parameters and sublists in the order they were added
    std::map<std::string, Entry> e_;
    std::map<std::string, ParameterList> sub_;
    template <class T> struct Tag {};
    static Entry make(int v) { Entry e{}; e.type = Entry::INT; e.i = v; return e; }
    static Entry make(double v) { Entry e{}; e.type = Entry::DOUBLE; e.d = v; return e; }
    static Entry make(bool v) { Entry e{}; e.type = Entry::BOOL; e.b = v; return e; }
    static Entry make(const std::string& v) { Entry e{}; e.type = Entry::STRING; e.s = v; return e; }
    [[noreturn]] void badType(const std::string& n) const {      // Teuchos::Exceptions::InvalidParameterType
        throw std::invalid_argument("the parameter \"" + n + "\" in the list \"" + name_ + "\" does not have the type that is asked for");
    }
    int& ref(Entry& e, const std::string& n, Tag<int>) { if (e.type != Entry::INT) badType(n); return e.i; }
    double& ref(Entry& e, const std::string& n, Tag<double>) { if (e.type != Entry::DOUBLE) badType(n); return e.d; }
    bool& ref(Entry& e, const std::string& n, Tag<bool>) { if (e.type != Entry::BOOL) badType(n); return e.b; }
    std::string& ref(Entry& e, const std::string& n, Tag<std::string>) { if (e.type != Entry::STRING) badType(n); return e.s; }

public:
    explicit ParameterList(const std::string& name = "ANONYMOUS") : name_(name) {}
    const std::string& name() const { return name_; }
    template <class T> ParameterList& set(const std::string& n, T v) {
        if (!e_.count(n)) order_.push_back(n);
        e_[n] = make(v);
        return *this;
    }
    ParameterList& set(const std::string& n, const char* v) { return set(n, std::string(v)); }
    // get(name, default): the entry, created with the default when it is missing (Teuchos' behaviour)
    template <class T> T& get(const std::string& n, T def) {
        auto it = e_.find(n);
        if (it == e_.end()) { order_.push_back(n); it = e_.emplace(n, make(def)).first; }
        return ref(it->second, n, Tag<T>());
    }
    template <class T> T& get(const std::string& n, const char* def) { return get<T>(n, std::string(def)); }
    template <class T> T& get(const std::string& n) {      // Teuchos::Exceptions::InvalidParameterName when it is missing
        auto it = e_.find(n);
        if (it == e_.end()) throw std::invalid_argument("the parameter \"" + n + "\" does not exist in the list \"" + name_ + "\"");
        return ref(it->second, n, Tag<T>());
    }
    ParameterList& sublist(const std::string& n) {
        auto it = sub_.find(n);
        if (it == sub_.end()) { order_.push_back(n); it = sub_.emplace(n, ParameterList(n)).first; }
        return it->second;
    }
    bool isParameter(const std::string& n) const { return e_.count(n) != 0; }
    bool isSublist(const std::string& n) const { return sub_.count(n) != 0; }
    const std::vector<std::string>& names() const { return order_; }
    const Entry& entry(const std::string& n) const { return e_.at(n); }
    const ParameterList& sublistConst(const std::string& n) const { return sub_.at(n); }
    // canonical text form, one "path = type value" line per parameter in insertion order (tests compare it with the Python reader's)
    void dump(std::ostream& os, const std::string& prefix = "") const {
        for (const auto& n : order_) {
            if (isSublist(n)) { sub_.at(n).dump(os, prefix + n + "/"); continue; }
            const Entry& e = e_.at(n);
            os << prefix << n << " = ";
            switch (e.type) {
            case Entry::INT: os << "int " << e.i; break;
            case Entry::DOUBLE: { char buf[64]; snprintf(buf, sizeof buf, "%.17g", e.d); os << "double " << buf; break; }
            case Entry::BOOL: os << "bool " << (e.b ? "true" : "false"); break;
            case Entry::STRING: os << "string " << e.s; break;
            }
            os << "\n";
        }
    }
};

namespace xml_detail {
inline std::string decode(const std::string& s) {
    static const std::pair<const char*, char> ent[] = {{"&amp;", '&'}, {"&lt;", '<'}, {"&gt;", '>'}, {"&quot;", '"'}, {"&apos;", '\''}};
    std::string o;
    for (size_t i = 0; i < s.size();) {
        bool hit = false;
        if (s[i] == '&')
            for (const auto& e : ent) { const size_t l = strlen(e.first); if (s.compare(i, l, e.first) == 0) { o += e.second; i += l; hit = true; break; } }
        if (!hit) o += s[i++];
    }
    return o;
}
inline std::string lower(std::string s) { for (auto& c : s) c = (char)tolower((unsigned char)c); return s; }
// attributes of one tag body: name="value" pairs, single or double quotes
inline std::map<std::string, std::string> attributes(const std::string& body, size_t pos) {
    std::map<std::string, std::string> a;
    while (pos < body.size()) {
        while (pos < body.size() && (isspace((unsigned char)body[pos]) || body[pos] == '/')) pos++;
        size_t eq = body.find('=', pos);
        if (eq == std::string::npos) break;
        std::string key = body.substr(pos, eq - pos);
        while (!key.empty() && isspace((unsigned char)key.back())) key.pop_back();
        size_t q = eq + 1;
        while (q < body.size() && isspace((unsigned char)body[q])) q++;
        if (q >= body.size() || (body[q] != '"' && body[q] != '\'')) throw std::invalid_argument("XML: attribute value without quotes in <" + body + ">");
        const size_t end = body.find(body[q], q + 1);
        if (end == std::string::npos) throw std::invalid_argument("XML: unterminated attribute value in <" + body + ">");
        a[key] = decode(body.substr(q + 1, end - q - 1));
        pos = end + 1;
    }
    return a;
}
}  // namespace xml_detail

// Teuchos::updateParametersFromXmlString into an empty list.  Numbers are read the way `istringstream >> value` reads them (leading
// number, trailing characters ignored: the reference's test/ocean/continuation_params.xml holds value="1.0-2"); bools are
// true / false / 1 / 0.
inline ParameterList parameterListFromXMLString(const std::string& text) {
    using namespace xml_detail;
    std::vector<ParameterList*> stack;
    ParameterList root;
    bool haveRoot = false;
    size_t pos = 0;
    while (true) {
        const size_t lt = text.find('<', pos);
        if (lt == std::string::npos) break;
        if (text.compare(lt, 4, "<!--") == 0) {
            const size_t e = text.find("-->", lt + 4);
            if (e == std::string::npos) throw std::invalid_argument("XML: unterminated comment");
            pos = e + 3;
            continue;
        }
        const size_t gt = text.find('>', lt);
        if (gt == std::string::npos) throw std::invalid_argument("XML: unterminated tag");
        const std::string body = text.substr(lt + 1, gt - lt - 1);
        pos = gt + 1;
        if (body.empty() || body[0] == '?' || body[0] == '!') continue;                  // declaration / doctype
        if (body[0] == '/') {                                                            // closing tag
            if (body.compare(1, 13, "ParameterList") == 0) { if (stack.empty()) throw std::invalid_argument("XML: unbalanced </ParameterList>"); stack.pop_back(); }
            continue;
        }
        size_t ne = 0;
        while (ne < body.size() && !isspace((unsigned char)body[ne]) && body[ne] != '/') ne++;
        const std::string tag = body.substr(0, ne);
        const auto at = attributes(body, ne);
        const bool selfClosing = body.back() == '/';
        if (tag == "ParameterList") {
            const std::string nm = at.count("name") ? at.at("name") : "ANONYMOUS";
            if (!haveRoot) { root = ParameterList(nm); haveRoot = true; if (!selfClosing) stack.push_back(&root); }
            else {
                if (stack.empty()) throw std::invalid_argument("XML: a second root list");
                ParameterList& sl = stack.back()->sublist(nm);
                if (!selfClosing) stack.push_back(&sl);
            }
        } else if (tag == "Parameter") {
            if (stack.empty()) throw std::invalid_argument("XML: <Parameter> outside a <ParameterList>");
            if (!at.count("name") || !at.count("type") || !at.count("value")) throw std::invalid_argument("XML: a <Parameter> needs name, type and value");
            const std::string ty = lower(at.at("type")), &val = at.at("value"), &nm = at.at("name");
            char* endp = nullptr;
            if (ty == "bool") {
                const std::string v = lower(val);
                if (v == "true" || v == "1") stack.back()->set(nm, true);
                else if (v == "false" || v == "0") stack.back()->set(nm, false);
                else throw std::invalid_argument("XML: cannot read \"" + val + "\" as bool (" + nm + ")");
            } else if (ty == "int" || ty == "long" || ty == "short" || ty == "unsigned int" || ty == "long long") {
                const long v = strtol(val.c_str(), &endp, 10);
                if (endp == val.c_str()) throw std::invalid_argument("XML: cannot read \"" + val + "\" as int (" + nm + ")");
                stack.back()->set(nm, (int)v);
            } else if (ty == "double" || ty == "float") {
                const double v = strtod(val.c_str(), &endp);
                if (endp == val.c_str()) throw std::invalid_argument("XML: cannot read \"" + val + "\" as double (" + nm + ")");
                stack.back()->set(nm, v);
            } else if (ty == "string" || ty == "char") stack.back()->set(nm, val);
            else throw std::invalid_argument("XML: unsupported parameter type \"" + at.at("type") + "\" (" + nm + ")");
        }
        // anything else (e.g. <Validators>) is ignored, as Teuchos does
    }
    if (!haveRoot) throw std::invalid_argument("XML: no <ParameterList> found");
    return root;
}
inline ParameterList parameterListFromXMLFile(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::invalid_argument("cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parameterListFromXMLString(ss.str());
}

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

// Ocean::Ocean(comm, Teuchos::ParameterList&) (Ocean.C:71-200, 985-1012): `oceanParams` = the list of ocean_params.xml with its "THCM"
// sublist and, as the reference's drivers merge it in, "Belos Solver" (solver_params.xml; defaults of Ocean::getDefaultInitParameters).
// "FGMRES iterations" is Belos' Num Blocks (the basis length), the iteration limit is Num Blocks x (Maximum Restarts + 1).
template <class ParameterList>
std::shared_ptr<Ocean> makeOcean(ParameterList& oceanParams, int rank = 0, int nranks = 1, int device = 0, int balance = 0) {
    auto& bs = oceanParams.sublist("Belos Solver");
    SolverParameters sp;
    sp.restart = bs.template get<int>("FGMRES iterations", 500);
    sp.tol = bs.template get<double>("FGMRES tolerance", 1e-8);
    sp.maxit = sp.restart * (bs.template get<int>("FGMRES restarts", 0) + 1);
    sp.precon = 1; sp.dgks = true;                                     // Belos "Orthogonalization" = "DGKS" (Ocean.C:1004)
    const THCMSetup su = setupFromParameterList(oceanParams.sublist("THCM"), rank, nranks, device, balance);
    return std::make_shared<Ocean>(makeTHCM(su), sp);
}

}  // namespace thcm_b200
