// Host-only test of include/thcm_paramlist.hpp (test infrastructure): the THCM constructor's reading of its parameter list over
// thcm_b200::ParameterList (Teuchos's accessor signatures: get<T>(name, default) stores the default when the entry is missing,
// sublist(name) creates the sublist) and the reader of the reference's XML dialect.  Runs without a GPU: only the m_global symbols of the
// library are called.
//   test_paramlist <directory that holds mkmask/> [<fixture directory>]   prints "PASS <n>" or the first failed check
//   test_paramlist --dump <file.xml>                                      prints the list in canonical form (compared with the Python reader)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include "thcm_paramlist.hpp"

using thcm_b200::ParameterList;

static int checks = 0;
#define CHECK(c) do { checks++; if (!(c)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
    using namespace thcm_b200;
    if (argc > 2 && std::string(argv[1]) == "--dump") {
        try { parameterListFromXMLFile(argv[2]).dump(std::cout); } catch (const std::exception& e) { printf("ERROR %s\n", e.what()); return 2; }
        return 0;
    }
    if (argc > 1) setenv("THCM_DATA_DIR", argv[1], 1);
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
    {   // the XML dialect: comments, nested lists, bools as 0 / 1 / true / false, numbers read like istringstream, entities
        ParameterList p = parameterListFromXMLString(
            "<?xml version=\"1.0\"?>\n<!-- a comment with <tags> inside -->\n<ParameterList name=\"Ocean\">\n"
            "  <Parameter name=\"Load state\" type=\"bool\" value=\"0\"/>\n  <Parameter value='1.0-2' type='double' name='odd number'/>\n"
            "  <ParameterList name=\"THCM\">\n    <Parameter name=\"Global Grid-Size n\" type=\"int\" value=\" 8 \"/>\n"
            "    <Parameter name=\"Land Mask\" type=\"string\" value=\"a &amp; b\"/>\n    <Parameter name=\"Periodic\" type=\"bool\" value=\"TRUE\"/>\n"
            "    <ParameterList name=\"Starting Parameters\">\n      <Parameter name=\"SPL1\" type=\"double\" value=\"2.0e3\"/>\n    </ParameterList>\n"
            "  </ParameterList>\n  <ParameterList name=\"Empty\"/>\n</ParameterList>\n");
        CHECK(p.name() == "Ocean" && p.get<bool>("Load state") == false && p.get<double>("odd number") == 1.0);
        CHECK(p.isSublist("THCM") && p.isSublist("Empty") && !p.isParameter("THCM"));
        ParameterList& t = p.sublist("THCM");
        CHECK(t.get<int>("Global Grid-Size n") == 8 && t.get<std::string>("Land Mask") == "a & b" && t.get<bool>("Periodic") == true);
        CHECK(t.sublist("Starting Parameters").get<double>("SPL1") == 2000.0);
        bool thrown = false;
        try { t.get<double>("Global Grid-Size n"); } catch (const std::invalid_argument&) { thrown = true; }     // wrong type
        CHECK(thrown);
        thrown = false;
        try { t.get<int>("no such parameter"); } catch (const std::invalid_argument&) { thrown = true; }          // missing name
        CHECK(thrown);
        thrown = false;
        try { parameterListFromXMLString("<ParameterList name='x'><Parameter name='a' type='bool' value='maybe'/></ParameterList>"); }
        catch (const std::invalid_argument&) { thrown = true; }
        CHECK(thrown);
    }
    if (argc > 2) {   // the committed fixtures through the XML reader and the constructor's rules
        const std::string dir = argv[2];
        ParameterList o = parameterListFromXMLFile(dir + "/natl8_integral_condition.xml");
        THCMSetup su = setupFromParameterList(o.sublist("THCM"));
        CHECK(su.settings.N == 8 && su.settings.L == 4 && su.settings.SRES == 0 && su.settings.vmix == 1 && su.settings.rho_mixing == 0);
        CHECK(su.integralCondition && su.Nic == 7 && su.Mic == 7 && su.startingParameters.size() == 5);
        CHECK(o.sublist("Belos Solver").get<int>("FGMRES iterations") == 120 && o.sublist("Belos Solver").get<double>("FGMRES tolerance") == 1e-6);
        ParameterList b = parameterListFromXMLFile(dir + "/basin16_topography1.xml");
        THCMSetup sb = setupFromParameterList(b);
        CHECK(sb.settings.N == 16 && sb.settings.forcing_type == 2 && sb.settings.SRES == 1 && !sb.integralCondition && sb.startingParameters.size() == 7);
        size_t ocean = 0;
        for (int v : sb.landm) ocean += v == 0;
        CHECK(ocean == 16u * 16 * 16);
    }
    printf("PASS %d\n", checks);
    return 0;
}
