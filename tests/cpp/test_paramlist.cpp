// Host-only test of include/thcm_paramlist.hpp (test infrastructure): the THCM constructor's reading of its parameter list over a small
// stand-in for Teuchos::ParameterList (Teuchos is not available here) -- same accessor signatures: get<T>(name, default) stores the
// default when the entry is missing, sublist(name) creates the sublist.  Runs without a GPU: only the m_global symbols of the library are
// called.  Usage: test_paramlist <directory that holds mkmask/>  ; prints "PASS <n>" or the first failed check.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include "thcm_paramlist.hpp"

class ParameterList {   // typed entries, Teuchos-style accessors
    struct Entry { enum { I, D, B, S } t; int i; double d; bool b; std::string s; };
    std::map<std::string, Entry> e_;
    std::map<std::string, ParameterList> sub_;
    template <class T> struct Tag {};
    static Entry make(int v) { Entry e{}; e.t = Entry::I; e.i = v; return e; }
    static Entry make(double v) { Entry e{}; e.t = Entry::D; e.d = v; return e; }
    static Entry make(bool v) { Entry e{}; e.t = Entry::B; e.b = v; return e; }
    static Entry make(const std::string& v) { Entry e{}; e.t = Entry::S; e.s = v; return e; }
    static int& ref(Entry& e, Tag<int>) { if (e.t != Entry::I) throw std::invalid_argument("type"); return e.i; }
    static double& ref(Entry& e, Tag<double>) { if (e.t != Entry::D) throw std::invalid_argument("type"); return e.d; }
    static bool& ref(Entry& e, Tag<bool>) { if (e.t != Entry::B) throw std::invalid_argument("type"); return e.b; }
    static std::string& ref(Entry& e, Tag<std::string>) { if (e.t != Entry::S) throw std::invalid_argument("type"); return e.s; }

public:
    template <class T> ParameterList& set(const std::string& name, T v) { e_[name] = make(v); return *this; }
    ParameterList& set(const std::string& name, const char* v) { e_[name] = make(std::string(v)); return *this; }
    template <class T> T& get(const std::string& name, T def) {
        auto it = e_.find(name);
        if (it == e_.end()) it = e_.emplace(name, make(def)).first;
        return ref(it->second, Tag<T>());
    }
    template <class T> T& get(const std::string& name, const char* def) { return get<T>(name, std::string(def)); }
    ParameterList& sublist(const std::string& name) { return sub_[name]; }
    bool isParameter(const std::string& name) const { return e_.count(name) != 0; }
};

static int checks = 0;
#define CHECK(c) do { checks++; if (!(c)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc > 1) setenv("THCM_DATA_DIR", argv[1], 1);
    using namespace thcm_b200;
    {   // all defaults: the reference's 16 x 16 x 16 North Atlantic box without continents (Topography = 1), restoring salinity
        ParameterList p;
        THCMSetup su = setupFromParameterList(p);
        const thcmb_settings& s = su.settings;
        CHECK(s.N == 16 && s.M == 16 && s.L == 16 && s.periodic == 0 && s.vmix == 1 && s.rho_mixing == 1 && s.SRES == 1 && s.TRES == 1);
        CHECK(std::fabs(s.xmin - 286.0 * 3.14159265358979323846 / 180.0) == 0.0 && s.hdim == 4000.0 && s.qz == 1.0 && s.forcing_type == 0);
        CHECK(!su.integralCondition && !su.fixPressurePoints && su.scaling == "THCM" && su.startingParameters.empty() && su.spert.empty());
        CHECK(p.isParameter("Mixing") && p.isParameter("Land Mask"));          // the defaults are written back, as Teuchos' get does
        size_t land = 0, ocean = 0;
        for (int v : su.landm) { land += v == 1; ocean += v == 0; }
        CHECK(su.landm.size() == 18u * 18 * 18 && ocean == 16u * 16 * 16 && land == su.landm.size() - ocean);
    }
    {   // the reference's 8 x 8 x 4 test box: mask by name, SRES = 0 -> integral condition at the default cell, starting parameters
        ParameterList p;
        p.set("Global Grid-Size n", 8).set("Global Grid-Size m", 8).set("Global Grid-Size l", 4).set("Read Land Mask", true)
            .set("Land Mask", "mask_natl8").set("Restoring Salinity Profile", 0).set("Rho Mixing", false).set("Fix Pressure Points", true);
        p.sublist("Starting Parameters").set("Combined Forcing", 0.25).set("SPL1", 2.0e3);
        THCMSetup su = setupFromParameterList(p, 1, 2, 1);
        CHECK(su.settings.rank == 1 && su.settings.nranks == 2 && su.settings.device == 1 && su.settings.rho_mixing == 0);
        CHECK(su.integralCondition && su.Nic == 7 && su.Mic == 7 && su.intSign == -1 && su.fixPressurePoints);
        CHECK(su.startingParameters.size() == 2 && su.startingParameters[0].first == "SPL1" && su.startingParameters[1].first == "Combined Forcing"
              && su.startingParameters[1].second == 0.25);                       // in the order of the parameter indices
        size_t land = 0;
        for (int k = 1; k <= 4; k++) for (int j = 1; j <= 8; j++) for (int i = 1; i <= 8; i++) land += su.landm[(size_t)i + 10 * (j + 10 * (size_t)k)] == 1;
        CHECK(land == 64);                                                       // mask_natl8: a quarter of the box is land
        // the rules of the constructor
        ParameterList q = p;
        q.set("Integral row coordinate i", 0).set("Integral row coordinate j", 0);
        bool thrown = false;
        try { setupFromParameterList(q); } catch (const std::invalid_argument& e) { thrown = std::strstr(e.what(), "land point") != nullptr; }
        CHECK(thrown);
        ParameterList r = p;
        r.set("Salinity Integral Sign", 3);
        thrown = false;
        try { setupFromParameterList(r); } catch (const std::invalid_argument&) { thrown = true; }
        CHECK(thrown);
        ParameterList c = p;
        c.set("Restoring Salinity Profile", 1).set("Coupled Salinity", 1);
        CHECK(setupFromParameterList(c).settings.SRES == 0);                     // THCM.C:253-259
        ParameterList t = p;
        t.set("Mixing", 2.0);                                                    // a double where an int is declared
        thrown = false;
        try { setupFromParameterList(t); } catch (const std::invalid_argument&) { thrown = true; }
        CHECK(thrown);
    }
    printf("PASS %d\n", checks);
    return 0;
}
