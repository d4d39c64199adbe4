// pc_ini.cpp -- the .ini driver path behind polychord_c_interface_ini.
//
// Replaces src/polychord/ini.f90 (read_params :44-95, get_string :149-224, get_params :354-458, get_prior_params
// :470-497) and the separable transforms of src/polychord/priors.f90 (uniform :40, gaussian :73, log_uniform :114,
// power_uniform :140, half_gaussian :155, exponential :166), their sorted forms (:242-360), the adaptive sorted
// families (:367-461) and nn_adaptive_layer_gaussian (:469-488).  Host-only.
#include "pc_ini.h"

#include <algorithm>
#include <cmath>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include "pc_errors.h"

#include "pc_device.cuh"  // inv_normal_cdf (AS241), host-callable

namespace pc {
namespace {

std::string trim(const std::string& s) {
    const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

// get_string (ini.f90:149-224): lines containing a comment character anywhere are skipped, the key is what stands
// before the first '=' or ':', the value what follows it; ith selects the ith match (1-based)
struct IniFile {
    std::vector<std::pair<std::string, std::string>> entries;
    explicit IniFile(const std::string& path) {
        std::ifstream f(path);
        if (!f) throw pc::ArgError("ini error: " + path + " does not exist");
        std::string line;
        while (std::getline(f, line)) {
            if (line.find_first_of("#!") != std::string::npos) continue;
            const size_t eq = line.find_first_of("=:");
            if (eq == std::string::npos) continue;
            entries.emplace_back(trim(line.substr(0, eq)), trim(line.substr(eq + 1)));
        }
    }
    std::string get(const std::string& key, int ith = 1) const {
        int c = 0;
        for (const auto& e : entries)
            if (e.first == key && ++c == ith) return e.second;
        return std::string();
    }
    int get_int(const std::string& key, int dflt, bool required = false) const {
        const std::string v = get(key);
        if (v.empty()) {
            if (required) throw pc::ArgError("ini error: '" + key + "' is missing");
            return dflt;
        }
        return std::stoi(v);
    }
    double get_double(const std::string& key, double dflt) const {
        const std::string v = get(key);
        return v.empty() ? dflt : std::stod(v);
    }
    bool get_logical(const std::string& key, bool dflt) const {  // Fortran list-directed logical: T / F / .true. / .false.
        std::string v = get(key);
        if (v.empty()) return dflt;
        size_t i = 0;
        if (v[0] == '.') i = 1;
        const char c = i < v.size() ? v[i] : 'f';
        return c == 'T' || c == 't';
    }
    std::vector<double> get_doubles(const std::string& key) const {
        std::vector<double> out;
        std::istringstream is(get(key));
        double v;
        while (is >> v) out.push_back(v);
        return out;
    }
};

std::vector<std::string> split_bar(const std::string& s) {
    std::vector<std::string> out;
    std::string cur;
    for (char c : s) {
        if (c == '|') { out.push_back(trim(cur)); cur.clear(); } else cur += c;
    }
    out.push_back(trim(cur));
    return out;
}

int prior_type_from_string(const std::string& s) {  // priors.f90:620-668
    static const char* names[] = {"", "uniform", "log_uniform", "power_uniform", "gaussian", "half_gaussian", "exponential",
                                  "sorted_uniform", "sorted_gaussian", "sorted_half_gaussian", "sorted_exponential",
                                  "adaptive_sorted_uniform", "adaptive_sorted_gaussian", "adaptive_sorted_half_gaussian",
                                  "adaptive_sorted_exponential", "nn_adaptive_layer_gaussian"};
    for (int i = 1; i < 16; ++i)
        if (s == names[i]) return i;
    return 0;
}

}  // namespace

IniConfig parse_ini(const std::string& path) {
    const IniFile f(path);
    IniConfig c;
    c.nlive = f.get_int("nlive", 0, true);                       // read_params, ini.f90:56-88
    c.num_repeats = f.get_int("num_repeats", 0, true);
    c.nprior = f.get_int("nprior", -1);
    c.nfail = f.get_int("nfail", -1);
    c.do_clustering = f.get_logical("do_clustering", false);
    c.feedback = f.get_int("feedback", 1);
    c.precision_criterion = f.get_double("precision_criterion", 1e-3);
    c.logzero = f.get_double("logzero", -1e30);
    c.max_ndead = f.get_int("max_ndead", -1);
    c.boost_posterior = f.get_double("boost_posterior", 0.0);
    c.posteriors = f.get_logical("posteriors", false);
    c.equals = f.get_logical("equals", false);
    c.cluster_posteriors = f.get_logical("cluster_posteriors", false);
    c.write_resume = f.get_logical("write_resume", false);
    c.write_paramnames = f.get_logical("write_paramnames", false);
    c.read_resume = f.get_logical("read_resume", false);
    c.write_stats = f.get_logical("write_stats", true);
    c.write_live = f.get_logical("write_live", false);
    c.write_dead = f.get_logical("write_dead", true);
    c.write_prior = f.get_logical("write_prior", false);
    c.maximise = f.get_logical("maximise", false);
    c.compression_factor = f.get_double("compression_factor", std::exp(-1.0));
    c.base_dir = f.get("base_dir").empty() ? "chains" : f.get("base_dir");
    c.file_root = f.get("file_root").empty() ? "test" : f.get("file_root");
    c.seed = f.get_int("seed", -1);
    c.grade_frac = f.get_doubles("grade_frac");
    if (c.grade_frac.empty()) c.grade_frac.assign(1, 1.0);
    if (!f.get("nlives").empty() || !f.get("loglikes").empty())
        throw pc::ArgError("ini error: dynamic nlive schedules (nlives / loglikes) are not supported by the B200 engine");
    // get_params, ini.f90:354-458:  P : name | latex | speed | prior type | prior block | prior params
    for (int i = 1;; ++i) {
        const std::string line = f.get("P", i);
        if (line.empty()) break;
        const std::vector<std::string> el = split_bar(line);
        if (el.size() < 6) throw pc::ArgError("ini error: parameter line needs 6 fields: " + line);
        IniParam p;
        p.name = el[0];
        const size_t star = p.name.find('*');   // sub-clustering marker (ini.f90:376): accepted, not used
        if (star != std::string::npos) p.name = p.name.substr(0, star);
        p.latex = el[1];
        p.speed = std::stoi(el[2]);
        p.prior_type = prior_type_from_string(el[3]);
        if (p.prior_type == 0) throw pc::ArgError("get_priors error: Unknown prior type for parameter " + p.name);
        p.block = std::stoi(el[4]);
        std::istringstream is(el[5]);
        double v;
        while (is >> v) p.params.push_back(v);
        const size_t need = p.prior_type == 3 ? 3 : ((p.prior_type == 6 || p.prior_type == 10 || p.prior_type == 14) ? 1 : 2);
        if (p.params.size() < need) throw pc::ArgError("ini error: too few prior parameters for " + p.name);
        c.params.push_back(p);
    }
    if (c.params.empty()) throw pc::ArgError("ini error: no parameters (P : ...) in " + path);
    for (int i = 1;; ++i) {
        const std::string line = f.get("D", i);
        if (line.empty()) break;
        const std::vector<std::string> el = split_bar(line);
        c.derived.push_back({el[0], el.size() > 1 ? el[1] : el[0]});
    }
    // grades: the parameters' speeds, slowest first (create_priors, priors.f90:671-749, orders them so; this driver
    // asks for them in that order)
    for (size_t i = 1; i < c.params.size(); ++i)
        if (c.params[i].speed < c.params[i - 1].speed)
            throw pc::ArgError("ini error: list the parameters in order of increasing speed (grade)");
    for (size_t i = 0; i < c.params.size(); ++i) {
        if (i == 0 || c.params[i].speed != c.params[i - 1].speed) c.grade_dims.push_back(0);
        c.grade_dims.back() += 1;
    }
    if (c.grade_frac.size() != c.grade_dims.size()) {
        if (c.grade_frac.size() == 1) c.grade_frac.assign(c.grade_dims.size(), c.grade_frac[0]);
        else throw pc::ArgError("ini error: grade_frac needs one entry per parameter speed");
    }
    return c;
}

static double separable_htp(int type, const double* q, double u) {
    switch (type) {
        case 1: return q[0] + (q[1] - q[0]) * u;                       // uniform_htp, priors.f90:40-55
        case 2: return q[0] * std::pow(q[1] / q[0], u);                 // log_uniform_htp, :114-124
        case 3: {                                                       // power_uniform_htp, :140-153
            const double a = std::pow(q[0], 1.0 / q[2]), b = std::pow(q[1], 1.0 / q[2]);
            return std::pow(a - u * std::fabs(a - b), q[2]);
        }
        case 4: return q[0] + q[1] * inv_normal_cdf(u);                 // gaussian_htp, :73-85
        case 5: return q[0] + q[1] * inv_normal_cdf(0.5 + 0.5 * u);     // half_gaussian_htp, :155-164
        default: return -std::log(1.0 - u) / q[0];                      // exponential_htp, :166-174
    }
}

// sort_hypercube (priors.f90:242-264): the unit cube onto its ordered corner, largest coordinate first
static void sort_hypercube(const double* u, double* s, size_t m) {
    if (m == 0) return;
    s[m - 1] = std::pow(u[m - 1], 1.0 / (double)m);
    for (size_t k = m - 1; k-- > 0;) s[k] = std::pow(u[k], 1.0 / (double)(k + 1)) * s[k + 1];
}

// adaptive_sorted_transform (priors.f90:367-385): the first coordinate, scaled to (0.5, m - 0.5), rounds to the number
// nfunc of basis functions in use; only the next nfunc coordinates are sorted, the rest pass through.  The reference
// indexes one past the block when the first coordinate is exactly 1; nfunc is clamped to m - 1 here.
static void adaptive_sorted_transform(const double* cube, double* t, size_t m) {
    for (size_t k = 0; k < m; ++k) t[k] = cube[k];
    if (m == 0) return;
    t[0] = 0.5 + cube[0] * (double)(m - 1);
    const size_t nfunc = std::min((size_t)(t[0] + 0.5), m - 1);
    sort_hypercube(cube + 1, t + 1, nfunc);
}

// adaptive_sorted_{uniform,gaussian,half_gaussian,exponential}_htp (priors.f90:389-461): the count coordinate is
// returned as scaled (its own prior parameters are skipped: "parameters(3:)", "(2:)" for the exponential), the others
// go through the separable transform of the same name
static void adaptive_block(const IniConfig& c, size_t i, size_t m, int base, const double* cube, double* theta) {
    std::vector<double> t(m);
    adaptive_sorted_transform(cube, t.data(), m);
    if (m) theta[0] = t[0];
    for (size_t k = 1; k < m; ++k) theta[k] = separable_htp(base, c.params[i + k].params.data(), t[k]);
}

void ini_prior_transform(const IniConfig& c, const double* cube, double* theta) {
    const size_t n = c.params.size();
    static const int base_type[] = {1, 4, 5, 6};   // uniform, gaussian, half_gaussian, exponential
    for (size_t i = 0; i < n;) {
        const IniParam& p = c.params[i];
        if (p.prior_type <= 6) {
            theta[i] = separable_htp(p.prior_type, p.params.data(), cube[i]);
            ++i;
            continue;
        }
        // block families: consecutive parameters of one type and one prior block (create_priors, priors.f90:671-749)
        size_t j = i;
        while (j < n && c.params[j].prior_type == p.prior_type && c.params[j].block == p.block) ++j;
        const size_t m = j - i;
        if (p.prior_type <= 10) {
            // sorted families (priors.f90:266-360): sort_hypercube, then the separable transform of the same name
            std::vector<double> srt(m);
            sort_hypercube(cube + i, srt.data(), m);
            for (size_t k = 0; k < m; ++k)
                theta[i + k] = separable_htp(base_type[p.prior_type - 7], c.params[i + k].params.data(), srt[k]);
        } else if (p.prior_type <= 14) {
            adaptive_block(c, i, m, base_type[p.prior_type - 11], cube + i, theta + i);
        } else {
            // nn_adaptive_layer_gaussian_htp (priors.f90:469-488): the first coordinate, scaled to (0.5, 2.5), is the
            // number of hidden layers; one layer -> adaptive sorted half-Gaussian on the rest, else adaptive sorted
            // Gaussian
            theta[i] = 0.5 + 2.0 * cube[i];
            if (m > 1) adaptive_block(c, i + 1, m - 1, theta[i] < 1.5 ? 5 : 4, cube + i + 1, theta + i + 1);
        }
        i = j;
    }
}

}  // namespace pc
