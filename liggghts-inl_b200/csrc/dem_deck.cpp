// dem_deck.cpp -- input-script front end of the B200 DEM engine: the reference's `lammps_open / lammps_command /
// lammps_file / lammps_close` (src/library.cpp:70-180) and the command grammar of src/input.cpp for the commands that
// configure the particle hot path.  Pure host C++ on top of the PUBLIC C ABI only (include/dem_b200.h):
// every deck command becomes the ABI call that replaces the reference object it would have created.
//
//   Input::file / Input::one / Input::parse / Input::substitute     src/input.cpp:160-560
//   read_data (header, Atoms, Velocities of atom_style sphere)        src/read_data.cpp:120-560, atom_vec_sphere.cpp:980-1060
//   region block, create_box                                           src/region_block.cpp:40-110, create_box.cpp:40-130
//   group id|type                                                      src/group.cpp:90-260
//   fix mesh/surface file ... [move|rotate|scale]                      src/fix_mesh.cpp:95-240,600-690, input_mesh_tri.cpp:308-591
//   run N [upto]                                                       src/run.cpp:40-130
//   fix particletemplate/sphere, particledistribution/discrete,        src/fix_template_sphere.cpp:72-366, fix_particledistribution_discrete.cpp:75-470,
//   fix insert/pack, region cylinder (host side of SURVEY.md 8f-3)     fix_insert.cpp:79-1030, fix_insert_pack.cpp:74-597, region.cpp:498-700, random_park.cpp
// Commands that only produce output (thermo, dump, compute, ...) are accepted and listed by dem_deck_warnings();
// anything else that is valid reference syntax but outside the hot path returns DEM_ERR_UNSUPPORTED.
//
// The same source builds the test-only oracle binding when DEM_DECK_ORACLE is defined (tests/: the entry points are then
// named orc_deck_* and drive the CPU oracle's orc_* functions) so that the parser is covered without a GPU.
#include <cctype>
#include <cmath>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#ifdef DEM_DECK_ORACLE
extern "C" {
typedef struct orc_engine dem_engine;
#define API(n) orc_##n
#define DECK(n) orc_deck_##n
#else
#include "../../include/dem_b200.h"
#define API(n) dem_##n
#define DECK(n) dem_deck_##n
extern "C" {
#endif
#ifdef DEM_DECK_ORACLE
const char *API(last_error)(const dem_engine *e);
int API(set_units)(dem_engine *e, const char *style);
int API(set_box)(dem_engine *e, const double lo[3], const double hi[3], const int periodic[3]);
int API(set_ntypes)(dem_engine *e, int ntypes);
int API(set_processors)(dem_engine *e, int px, int py, int pz);
int API(set_neighbor)(dem_engine *e, double skin, int every, int delay, int check);
int API(set_timestep)(dem_engine *e, double dt);
int API(set_contact_distance_factor)(dem_engine *e, double f);
int API(set_property)(dem_engine *e, const char *name, const char *kind, const double *values, int n);
int API(set_pair_style)(dem_engine *e, int argc, const char *const *argv);
int API(add_wall_primitive)(dem_engine *e, const char *id, int argc, const char *const *argv);
int API(add_mesh)(dem_engine *e, const char *id, int atom_type, const double *nodes9, long ntri, int argc, const char *const *argv);
int API(move_mesh)(dem_engine *e, const char *mesh_id, int argc, const char *const *argv);
int API(add_wall_mesh)(dem_engine *e, const char *id, int argc, const char *const *argv);
int API(set_gravity)(dem_engine *e, double magnitude, const double dir[3]);
int API(set_freeze)(dem_engine *e, int groupbit);
int API(set_integrate)(dem_engine *e, int groupbit);
int API(set_extra_force)(dem_engine *e, const char *id, int kind, int groupbit, const double *values, int n);
int API(upload_particles)(dem_engine *e, long n, const int *tag, const int *type, const int *mask, const double *x, const double *v,
                          const double *omega, const double *radius, const double *density);
int API(insert_particles)(dem_engine *e, long n, const int *tag, const int *type, const int *mask, const double *x, const double *v,
                          const double *omega, const double *radius, const double *density);
int API(insert_step_begin)(dem_engine *e);
int API(insert_step_end)(dem_engine *e, long n, const int *tag, const int *type, const int *mask, const double *x, const double *v,
                         const double *omega, const double *radius, const double *density);
long API(nlocal)(const dem_engine *e);
int API(download)(dem_engine *e, const char *field, void *out, long count);
int API(setup)(dem_engine *e);
int API(run)(dem_engine *e, long nsteps);
#endif
}

namespace {

enum { OK = 0, ERR_ARG = -1, ERR_UNSUPPORTED = -2, ERR_STATE = -3 };

struct Deck {
  dem_engine *e = nullptr;
  std::string err, warnings, dir;  // dir: directory of the deck file (relative paths in read_data / mesh files)
  std::map<std::string, std::string> vars;
  std::set<std::string> var_is_equal;  // equal-style: the stored text is a formula
  // Park-Miller minimal standard generator, the stream every insertion object of the reference draws from (random_park.cpp:72-110)
  struct Park {
    int seed = 1, save = 0; double second = 0.0;
    double uniform() { const int k = seed / 127773; seed = 16807 * (seed - k * 127773) - 2836 * k; if (seed < 0) seed += 2147483647; return (1.0 / 2147483647) * seed; }
    double gaussian() {
      if (save) { save = 0; return second; }
      double v1, v2, rsq;
      do { v1 = 2.0 * uniform() - 1.0; v2 = 2.0 * uniform() - 1.0; rsq = v1 * v1 + v2 * v2; } while (!(rsq < 1.0 && rsq != 0.0));
      const double fac = std::sqrt(-2.0 * std::log(rsq) / rsq);
      second = v1 * fac; save = 1; return v2 * fac;
    }
    void reset(int s) { seed = s; save = 0; }
  };
  struct Region {
    int kind = 0;  // 0 block (lo, hi), 1 cylinder (axis, c1, c2, rad, clo, chi)
    double lo[3], hi[3];
    char axis = 'z'; double c1 = 0, c2 = 0, rad = 0, clo = 0, chi = 0;
    double ext_lo[3], ext_hi[3];  // bounding box (Region::extent_*)
    Park random;                  // region.cpp:456-457
  };
  std::map<std::string, Region> regions;
  // fix particletemplate/sphere, fix particledistribution/discrete, fix insert/pack (host side of the insertion: where and when;
  // the particles themselves enter the engine through dem_insert_step_begin / _end)
  struct Template { int atom_type = 1, groupbit = 1; double radius = 0, density = 0, volexpect = 0, massexpect = 0; };
  struct Distribution { Park random; int groupbit = 1; std::vector<std::string> templates; std::vector<double> weight; std::vector<int> order; double volexpect = 0, massexpect = 0, maxrbound = 0; };
  struct Insert {
    std::string id, dist, region; int groupbit = 1, seed = 0; Park random;
    int vmode = 0; double v[3] = {0, 0, 0}, vfluct[3] = {0, 0, 0}, omega[3] = {0, 0, 0};
    long insert_every = -1, next = 0; int maxattempt = 50, check_ol = 1, all_in = 0, exact_number = 1, ntry_mc = 100000;
    double volumefraction = 0.0, masstotal = 0.0; long ntotal = 0;
    bool setup_done = false, check_border = true, warn_region = true;
    double region_volume = 0.0, region_volume_local = 0.0, insertion_ratio = 0.0;
    long ninserted = 0; double massinserted = 0.0;
  };
  // lattice (lattice.cpp:60-300): styles sc | bcc | fcc with the default orientation; scale = lattice constant; optional origin
  struct Lattice { bool defined = false; double scale = 0.0, origin[3] = {0, 0, 0}; std::vector<double> basis; } lattice;
  std::map<std::string, Template> templates;
  std::map<std::string, Distribution> distributions;
  std::vector<Insert> inserts;  // in the order of definition (== the order Modify calls pre_exchange)
  int nprocs = 1;
  std::set<std::string> opaque_regions;  // regions of other shapes: known by name only
  std::map<std::string, int> groups;  // name -> mask bit
  std::map<std::string, std::string> ignored_fixes;
  bool have_freeze = false;
  std::set<std::string> extra_fixes;  // fix addforce / viscous ids (dem_set_extra_force)
  // box / particles held until the first `run`
  bool have_box = false, box_sent = false, uploaded = false, newton_off = false, have_pair = false;
  double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  int periodic[3] = {0, 0, 0}, ntypes = 0;
  double skin = 0.0; int every = 1, delay = 10, check = 1;  // neighbor.cpp:100-130 defaults (delay 10, every 1, check yes)
  std::vector<int> tag, type, mask;
  std::vector<double> x, v, omega, radius, density;
  // atoms created after the first run (create_atoms single + set atom ...): handed to dem_insert_particles by the next run
  std::vector<int> ptag, ptype, pmask;
  std::vector<double> px, pv, pomega, pradius, pdensity;
  int maxtag = 0;
  // dump custom (dump_custom.cpp): text snapshots every N steps, atoms in ascending id
  struct Dump { std::string file; long every = 0, last = -1; std::vector<std::string> fields; int pad = 0; bool first = false; };  // (dump_modify pad / first, dump.cpp:640-700)
  std::map<std::string, Dump> dumps;
  std::string bstr[3] = {"ff", "ff", "ff"};  // Domain::boundary_string
  // thermo (thermo.cpp): one line at the start of a run, on the multiples of N and on the last step
  long thermo_every = 0; double dt = 0.0;
  std::vector<std::string> thermo_kw = {"step", "atoms", "ke", "cpu"};  // thermo.cpp:98 (style one)
  bool screen = false;      // print the lines to stdout as they come (lmp_b200); they are always kept in `out`
  std::string out;
  std::chrono::steady_clock::time_point loop0;
  long ntimestep = 0;
};

int fail(Deck *d, int code, const char *fmt, ...)
{
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  d->err = buf;
  return code;
}
// an ABI call failed: keep its status, take the engine's message
int engine_failed(Deck *d, int rc) { d->err = API(last_error)(d->e); return rc; }
#define TRY(call) do { const int rc_ = (call); if (rc_ != 0) return engine_failed(d, rc_); } while (0)

bool is_number(const std::string &s) { if (s.empty()) return false; char *end = nullptr; strtod(s.c_str(), &end); return end && *end == 0; }
// Force::numeric / inumeric (force.cpp:740-800): the whole token must be a number
int numeric(Deck *d, const std::string &s, double &out)
{
  if (!is_number(s)) return fail(d, ERR_ARG, "Expected floating point parameter in input script or data file: '%s'", s.c_str());
  out = atof(s.c_str()); return OK;
}
int inumeric(Deck *d, const std::string &s, int &out)
{
  if (s.empty()) return fail(d, ERR_ARG, "Expected integer parameter in input script or data file");
  for (size_t k = 0; k < s.size(); k++) if (!(isdigit((unsigned char)s[k]) || (k == 0 && (s[k] == '-' || s[k] == '+')))) return fail(d, ERR_ARG, "Expected integer parameter in input script or data file: '%s'", s.c_str());
  out = atoi(s.c_str()); return OK;
}

// equal-style variable formulas (variable.cpp:700-1500): numbers, + - * / ^, unary minus, parentheses, v_name references to
// other variables, PI, and the math functions sqrt exp ln log abs sin cos tan asin acos atan floor ceil round.  Anything else
// (thermo keywords, atom values vx[1], compute / fix references ...) only feeds output commands and is reported as
// unsupported when -- and only when -- a hot-path command asks for the value.
struct Formula {
  Deck *d; const char *p; std::string err; int depth;
  void ws() { while (*p == ' ' || *p == '\t') p++; }
  double fail(const std::string &m) { if (err.empty()) err = m; return 0.0; }
  double atom()
  {
    ws();
    if (*p == '(') { p++; const double v = expr(); ws(); if (*p != ')') return fail("Invalid syntax in variable formula"); p++; return v; }
    // Variable::evaluate precedence (variable.cpp:133-141): UNARY (8) binds tighter than CARAT (7), so -2^2 == 4
    if (*p == '-') { p++; return -atom(); }
    if (*p == '+') { p++; return atom(); }
    if (isdigit((unsigned char)*p) || *p == '.') { char *e; const double v = strtod(p, &e); p = e; return v; }
    if (isalpha((unsigned char)*p) || *p == '_') {
      std::string id; while (isalnum((unsigned char)*p) || *p == '_') id += *p++;
      ws();
      if (id == "PI") return 3.14159265358979323846;
      if (id.compare(0, 2, "v_") == 0) return value_of(id.substr(2));
      if (*p == '(') {
        p++; const double a = expr(); ws(); if (*p != ')') return fail("Invalid syntax in variable formula"); p++;
        if (id == "sqrt") return a < 0 ? fail("Sqrt of negative value in variable formula") : std::sqrt(a);
        if (id == "exp") return std::exp(a);
        if (id == "ln") return a <= 0 ? fail("Log of zero/negative value in variable formula") : std::log(a);
        if (id == "log") return a <= 0 ? fail("Log of zero/negative value in variable formula") : std::log10(a);
        if (id == "abs") return std::fabs(a);
        if (id == "sin") return std::sin(a);
        if (id == "cos") return std::cos(a);
        if (id == "tan") return std::tan(a);
        if (id == "asin") return std::asin(a);
        if (id == "acos") return std::acos(a);
        if (id == "atan") return std::atan(a);
        if (id == "floor") return std::floor(a);
        if (id == "ceil") return std::ceil(a);
        if (id == "round") return std::floor(a + 0.5);
      }
      return fail("variable formula term '" + id + "' is outside the hot-path scope");
    }
    return fail("Invalid syntax in variable formula");
  }
  // '^' pops on >= precedence in the reference, i.e. it is LEFT-associative: 2^3^2 == 64
  double power() { double b = atom(); for (;;) { ws(); if (*p == '^') { p++; const double e = atom(); b = std::pow(b, e); } else return b; } }
  double term()
  {
    double v = power();
    for (;;) { ws(); if (*p == '*') { p++; v *= power(); } else if (*p == '/') { p++; const double q = power(); if (q == 0.0) return fail("Divide by 0 in variable formula"); v /= q; } else return v; }
  }
  double expr() { double v = term(); for (;;) { ws(); if (*p == '+') { p++; v += term(); } else if (*p == '-') { p++; v -= term(); } else return v; } }
  double value_of(const std::string &name);
};
int evaluate(Deck *d, const std::string &text, double &out, int depth);
double Formula::value_of(const std::string &name)
{
  auto it = d->vars.find(name);
  if (it == d->vars.end()) return fail("Invalid variable name '" + name + "' in variable formula");
  if (depth > 32) return fail("variable formulas reference each other in a loop");
  double v = 0.0;
  const bool formula = d->var_is_equal.count(name) != 0;
  if (formula) { if (evaluate(d, it->second, v, depth + 1) != OK) return fail(d->err); return v; }
  if (!is_number(it->second)) return fail("Variable '" + name + "' does not hold a number");
  return atof(it->second.c_str());
}
int evaluate(Deck *d, const std::string &text, double &out, int depth = 0)
{
  Formula f{d, text.c_str(), "", depth};
  out = f.expr(); f.ws();
  if (f.err.empty() && *f.p) f.err = "Invalid syntax in variable formula";
  if (!f.err.empty()) return fail(d, f.err.find("outside the hot-path scope") != std::string::npos ? ERR_UNSUPPORTED : ERR_ARG, "%s: '%s'", f.err.c_str(), text.c_str());
  return OK;
}

// Input::substitute (input.cpp:430-520): ${name} and $x
int substitute(Deck *d, std::string &line)
{
  std::string out; bool quote = false;
  for (size_t k = 0; k < line.size(); k++) {
    const char c = line[k];
    if (c == '"') quote = !quote;
    if (c == '$' && !quote && k + 1 < line.size()) {
      std::string name; size_t end;
      if (line[k + 1] == '{') { end = line.find('}', k + 2); if (end == std::string::npos) return fail(d, ERR_ARG, "Invalid variable name"); name = line.substr(k + 2, end - k - 2); }
      else { name = line.substr(k + 1, 1); end = k + 1; }
      auto it = d->vars.find(name);
      if (it == d->vars.end()) return fail(d, ERR_ARG, "Substitution for illegal variable '%s'", name.c_str());
      if (d->var_is_equal.count(name)) {  // Variable::retrieve, variable.cpp:595-606: the formula is evaluated now, printed %.15g
        double v; const int rc = evaluate(d, it->second, v); if (rc) return rc;
        char buf[64]; snprintf(buf, sizeof buf, "%.15g", v);
        out += buf;
      } else out += it->second;
      k = end; continue;
    }
    out += c;
  }
  line = out; return OK;
}
// Input::parse (input.cpp:330-420): '#' starts a comment outside quotes; words split on whitespace, "..." kept as one word
std::vector<std::string> split(std::string line)
{
  bool quote = false;
  for (size_t k = 0; k < line.size(); k++) { if (line[k] == '"') quote = !quote; if (line[k] == '#' && !quote) { line.resize(k); break; } }
  std::vector<std::string> w; std::string cur; quote = false; bool have = false;
  for (char c : line) {
    if (c == '"') { quote = !quote; have = true; continue; }
    if (!quote && isspace((unsigned char)c)) { if (have) { w.push_back(cur); cur.clear(); have = false; } continue; }
    cur += c; have = true;
  }
  if (have) w.push_back(cur);
  return w;
}
std::string resolve(const Deck *d, const std::string &path)
{
  if (path.empty() || path[0] == '/' || d->dir.empty()) return path;
  std::ifstream probe(path.c_str());
  return probe.good() ? path : d->dir + "/" + path;
}
std::vector<const char *> cptrs(const std::vector<std::string> &w, size_t from)
{
  std::vector<const char *> a;
  for (size_t k = from; k < w.size(); k++) a.push_back(w[k].c_str());
  return a;
}

// ---- read_data: header + Atoms + Velocities of atom_style sphere (id type diameter density x y z [ix iy iz]; id vx vy vz wx wy wz)
int read_data(Deck *d, const std::string &path)
{
  std::ifstream f(resolve(d, path).c_str());
  if (!f.good()) return fail(d, ERR_ARG, "Cannot open file %s", path.c_str());
  std::string line;
  std::getline(f, line);  // title
  long natoms = -1; int ntypes = -1; double lo[3], hi[3]; bool hb[3] = {false, false, false};
  std::string section;
  auto strip = [](std::string s) { const size_t h = s.find('#'); if (h != std::string::npos) s.resize(h); return s; };
  while (std::getline(f, line)) {
    const std::vector<std::string> w = split(strip(line));
    if (w.empty()) continue;
    if (w.size() == 2 && w[1] == "atoms") natoms = atol(w[0].c_str());
    else if (w.size() == 3 && w[1] == "atom" && w[2] == "types") ntypes = atoi(w[0].c_str());
    else if (w.size() == 4 && w[2] == "xlo" && w[3] == "xhi") { lo[0] = atof(w[0].c_str()); hi[0] = atof(w[1].c_str()); hb[0] = true; }
    else if (w.size() == 4 && w[2] == "ylo" && w[3] == "yhi") { lo[1] = atof(w[0].c_str()); hi[1] = atof(w[1].c_str()); hb[1] = true; }
    else if (w.size() == 4 && w[2] == "zlo" && w[3] == "zhi") { lo[2] = atof(w[0].c_str()); hi[2] = atof(w[1].c_str()); hb[2] = true; }
    else if (w.size() >= 2 && (w[1] == "bonds" || w[1] == "angles" || w[1] == "dihedrals" || w[1] == "impropers" || w[1] == "bond" || w[1] == "angle" || w[1] == "extra")) continue;
    else if (isalpha((unsigned char)w[0][0])) { section = w[0]; break; }
    else return fail(d, ERR_ARG, "Unknown identifier in data file: %s", line.c_str());
  }
  if (natoms < 0) return fail(d, ERR_ARG, "No atoms in data file");
  if (!d->have_box) {
    if (!(hb[0] && hb[1] && hb[2]) || ntypes < 1) return fail(d, ERR_ARG, "data file header needs 'atom types' and the box bounds");
    for (int k = 0; k < 3; k++) { d->lo[k] = lo[k]; d->hi[k] = hi[k]; }
    d->ntypes = ntypes; d->have_box = true;
  }
  const size_t base = d->tag.size();
  std::map<int, size_t> index;
  bool got_atoms = false;
  while (!section.empty()) {
    const std::string sec = section; section.clear();
    long nread = 0;
    if (sec == "Atoms") {
      d->tag.resize(base + natoms); d->type.resize(base + natoms); d->mask.resize(base + natoms, 1);
      d->x.resize(3 * (base + natoms)); d->v.resize(3 * (base + natoms), 0.0); d->omega.resize(3 * (base + natoms), 0.0);
      d->radius.resize(base + natoms); d->density.resize(base + natoms);
    } else if (sec != "Velocities") return fail(d, ERR_UNSUPPORTED, "data file section '%s' is outside the hot-path scope", sec.c_str());
    while (nread < natoms && std::getline(f, line)) {
      const std::vector<std::string> w = split(strip(line));
      if (w.empty()) continue;
      if (sec == "Atoms") {
        if (w.size() != 7 && w.size() != 10) return fail(d, ERR_ARG, "Incorrect atom format in data file");
        const size_t i = base + nread;
        d->tag[i] = atoi(w[0].c_str()); d->type[i] = atoi(w[1].c_str()); d->mask[i] = 1;
        d->radius[i] = 0.5 * atof(w[2].c_str());  // atom_vec_sphere.cpp data_atom: radius = 0.5 * diameter
        d->density[i] = atof(w[3].c_str());
        for (int k = 0; k < 3; k++) d->x[3 * i + k] = atof(w[4 + k].c_str());
        if (d->tag[i] <= 0) return fail(d, ERR_ARG, "Invalid atom ID in Atoms section of data file");
        if (d->type[i] <= 0 || d->type[i] > d->ntypes) return fail(d, ERR_ARG, "Invalid atom type in Atoms section of data file");
        if (!(d->radius[i] >= 0.0) || !(d->density[i] > 0.0)) return fail(d, ERR_ARG, "Invalid radius or density in Atoms section of data file");
        index[d->tag[i]] = i;
      } else {
        if (!got_atoms) return fail(d, ERR_ARG, "Must read Atoms before Velocities");
        if (w.size() != 7) return fail(d, ERR_ARG, "Incorrect velocity format in data file");
        auto it = index.find(atoi(w[0].c_str()));
        if (it == index.end()) return fail(d, ERR_ARG, "Invalid atom ID in Velocities section of data file");
        for (int k = 0; k < 3; k++) { d->v[3 * it->second + k] = atof(w[1 + k].c_str()); d->omega[3 * it->second + k] = atof(w[4 + k].c_str()); }
      }
      nread++;
    }
    if (nread != natoms) return fail(d, ERR_ARG, "Unexpected end of data file");
    if (sec == "Atoms") got_atoms = true;
    while (std::getline(f, line)) { const std::vector<std::string> w = split(strip(line)); if (!w.empty()) { section = w[0]; break; } }
  }
  if (!got_atoms) return fail(d, ERR_ARG, "No Atoms section in data file");
  return OK;
}

// ---- STL reader: InputMeshTri::meshtrifile_stl (ASCII, vertices by atof) and meshtrifile_stl_binary, input_mesh_tri.cpp:308-591
int read_stl(Deck *d, const std::string &path, std::vector<double> &nodes)
{
  const std::string p = resolve(d, path);
  std::ifstream f(p.c_str(), std::ios::binary);
  if (!f.good()) return fail(d, ERR_ARG, "Cannot open mesh file %s", path.c_str());
  std::string first;
  std::getline(f, first);
  const std::vector<std::string> fw = split(first);
  const bool ascii = !fw.empty() && fw[0].compare(0, 5, "solid") == 0;
  nodes.clear();
  if (ascii) {
    std::string line; int nv = 0; bool infacet = false;
    while (std::getline(f, line)) {
      const std::vector<std::string> w = split(line);
      if (w.empty()) continue;
      if (w[0] == "facet") { if (infacet) return fail(d, ERR_ARG, "Corrupt or unknown STL file: New facet begins without closing prior facet."); infacet = true; nv = 0; }
      else if (w[0] == "vertex") {
        if (!infacet || w.size() < 4) return fail(d, ERR_ARG, "Corrupt or unknown STL file: Vertex found outside a loop.");
        if (nv == 3) return fail(d, ERR_ARG, "Corrupt or unknown STL file: Found more than 3 vertices in a facet (only triangular meshes supported).");
        for (int k = 0; k < 3; k++) nodes.push_back(atof(w[1 + k].c_str()));
        nv++;
      } else if (w[0] == "endfacet") { if (!infacet || nv != 3) return fail(d, ERR_ARG, "Corrupt or unknown STL file: End of facet found, but no begin."); infacet = false; }
    }
  } else {
    f.clear(); f.seekg(80, std::ios::beg);
    unsigned int nf = 0;
    f.read((char *)&nf, 4);
    for (unsigned int t = 0; t < nf; t++) {
      float rec[12]; unsigned short attr;
      f.read((char *)rec, 48); f.read((char *)&attr, 2);
      if (!f.good()) return fail(d, ERR_ARG, "Corrupt STL file: Error in reading binary STL file.");
      for (int k = 3; k < 12; k++) nodes.push_back((double)rec[k]);
    }
  }
  if (nodes.empty()) return fail(d, ERR_ARG, "mesh file %s holds no triangles", path.c_str());
  return OK;
}
// load-time transforms of `fix mesh/surface` in keyword order: FixMesh::moveMesh / rotateMesh / scaleMesh (fix_mesh.cpp:600-690)
// -> MultiNodeMesh::move / rotate(dAngle, axis, p = 0) / scale (multi_node_mesh_I.h:470-700), same floating-point expressions
void quatquat(const double *a, const double *b, double *c)
{
  c[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  c[1] = a[0] * b[1] + b[0] * a[1] + a[2] * b[3] - a[3] * b[2];
  c[2] = a[0] * b[2] + b[0] * a[2] + a[3] * b[1] - a[1] * b[3];
  c[3] = a[0] * b[3] + b[0] * a[3] + a[1] * b[2] - a[2] * b[1];
}
void mesh_rotate(std::vector<double> &nodes, const double axis[3], double phi_deg)
{
  const double dAngle = phi_deg * 3.14159265 / 180.0;  // FixMesh::rotateMesh uses this literal (fix_mesh.cpp:656)
  double ax[3] = {axis[0], axis[1], axis[2]};
  const double sinv = 1. / std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
  for (int k = 0; k < 3; k++) ax[k] = sinv * ax[k];
  double q[4] = {std::cos(dAngle * 0.5), ax[0] * std::sin(dAngle * 0.5), ax[1] * std::sin(dAngle * 0.5), ax[2] * std::sin(dAngle * 0.5)};
  const double qc[4] = {q[0], -q[1], -q[2], -q[3]};
  for (size_t n = 0; n + 2 < nodes.size(); n += 3) {
    const double vq[4] = {0., nodes[n], nodes[n + 1], nodes[n + 2]};
    double t[4], r[4];
    quatquat(q, vq, t); quatquat(t, qc, r);
    nodes[n] = r[1]; nodes[n + 1] = r[2]; nodes[n + 2] = r[3];
  }
}

// thermo: header (thermo.cpp:311-330: "%8s " for integers, "%14s " for floats) and one line of values ("%8ld ", "%14.8g ")
struct ThermoCol { const char *kw, *head; bool is_int; };
static const ThermoCol thermo_cols[] = {{"step", "Step", true}, {"atoms", "Atoms", true}, {"ke", "KinEng", false}, {"erotate", "RotEng", false},
                                        {"cpu", "CPU", false}, {"time", "Time", false}, {"elapsed", "Elapsed", true}, {nullptr, nullptr, false}};
const ThermoCol *thermo_col(const std::string &kw) { for (const ThermoCol *c = thermo_cols; c->kw; c++) if (kw == c->kw) return c; return nullptr; }
void emit(Deck *d, const std::string &line) { d->out += line; if (d->screen) { fputs(line.c_str(), stdout); fflush(stdout); } }
int thermo_header(Deck *d)
{
  std::string line; char buf[64];
  for (auto &kw : d->thermo_kw) { const ThermoCol *c = thermo_col(kw); snprintf(buf, sizeof buf, c->is_int ? "%8s " : "%14s ", c->head); line += buf; }
  emit(d, line + "\n"); return OK;
}
int thermo_line(Deck *d, bool first, long run_first)
{
  const long n = API(nlocal)(d->e);
  double ke = 0.0, erot = 0.0;
  bool need_ke = false, need_er = false;
  for (auto &kw : d->thermo_kw) { need_ke |= kw == "ke"; need_er |= kw == "erotate"; }
  if ((need_ke || need_er) && n) {  // compute_ke.cpp:60-80, compute_erotate_sphere.cpp:60-90 (mvv2e = 1 in the supported unit systems)
    std::vector<double> m(n), v(3 * n);
    TRY(API(download)(d->e, "rmass", m.data(), n));
    if (need_ke) { TRY(API(download)(d->e, "v", v.data(), n)); for (long i = 0; i < n; i++) ke += m[i] * (v[3 * i] * v[3 * i] + v[3 * i + 1] * v[3 * i + 1] + v[3 * i + 2] * v[3 * i + 2]); ke *= 0.5; }
    if (need_er) {
      std::vector<double> r(n);
      TRY(API(download)(d->e, "omega", v.data(), n)); TRY(API(download)(d->e, "radius", r.data(), n));
      for (long i = 0; i < n; i++) erot += (v[3 * i] * v[3 * i] + v[3 * i + 1] * v[3 * i + 1] + v[3 * i + 2] * v[3 * i + 2]) * r[i] * r[i] * m[i];
      erot *= 0.5 * 0.4;
    }
  }
  const double cpu = first ? 0.0 : std::chrono::duration<double>(std::chrono::steady_clock::now() - d->loop0).count();
  std::string line; char buf[64];
  for (auto &kw : d->thermo_kw) {
    if (kw == "step") snprintf(buf, sizeof buf, "%8ld ", d->ntimestep);
    else if (kw == "atoms") snprintf(buf, sizeof buf, "%8ld ", n);
    else if (kw == "elapsed") snprintf(buf, sizeof buf, "%8ld ", d->ntimestep - run_first);
    else if (kw == "ke") snprintf(buf, sizeof buf, "%14.8g ", ke);
    else if (kw == "erotate") snprintf(buf, sizeof buf, "%14.8g ", erot);
    else if (kw == "cpu") snprintf(buf, sizeof buf, "%14.8g ", cpu);
    else if (kw == "time") snprintf(buf, sizeof buf, "%14.8g ", d->ntimestep * d->dt);
    line += buf;
  }
  emit(d, line + "\n"); return OK;
}

// dump custom: one snapshot (dump_custom.cpp:380-392 header, :1027-1042 lines: every value "%d " / "%g ", atoms in ascending id)
int write_dump(Deck *d, Deck::Dump &D)
{
  const long n = API(nlocal)(d->e);
  std::string path = D.file;
  const size_t star = path.find('*');
  const bool per_step = star != std::string::npos;
  if (per_step) { char num[48]; snprintf(num, sizeof num, "%0*ld", D.pad, d->ntimestep); path = path.substr(0, star) + num + path.substr(star + 1); }
  if (!path.empty() && path[0] != '/' && !d->dir.empty()) path = d->dir + "/" + path;
  FILE *fp = fopen(path.c_str(), (per_step || D.last < 0) ? "w" : "a");
  if (!fp) return fail(d, ERR_ARG, "Cannot open dump file %s", path.c_str());
  fprintf(fp, "ITEM: TIMESTEP\n%ld\nITEM: NUMBER OF ATOMS\n%ld\nITEM: BOX BOUNDS %s %s %s\n", d->ntimestep, n, d->bstr[0].c_str(), d->bstr[1].c_str(), d->bstr[2].c_str());
  for (int k = 0; k < 3; k++) fprintf(fp, "%g %g\n", d->lo[k], d->hi[k]);
  fprintf(fp, "ITEM: ATOMS");
  for (auto &f : D.fields) fprintf(fp, " %s", f.c_str());
  fprintf(fp, " \n");
  std::map<std::string, std::vector<double>> dv; std::map<std::string, std::vector<int>> iv;
  auto needd = [&](const char *f, int w) -> int { if (dv.count(f)) return OK; dv[f].resize((size_t)std::max<long>(n, 1) * w); return n ? API(download)(d->e, f, dv[f].data(), n) : OK; };
  auto needi = [&](const char *f) -> int { if (iv.count(f)) return OK; iv[f].resize((size_t)std::max<long>(n, 1)); return n ? API(download)(d->e, f, iv[f].data(), n) : OK; };
  struct Col { int kind; const void *p; int w, c; double scale; };  // kind 0 int, 1 double
  std::vector<Col> cols;
  for (auto &f : D.fields) {
    int rc = OK; Col c{1, nullptr, 1, 0, 1.0};
    auto vec3 = [&](const char *fld, const char *base) -> bool {
      const std::string b(base);
      for (int k = 0; k < 3; k++) if (f == b + "xyz"[k]) { rc = needd(fld, 3); c = Col{1, dv[fld].data(), 3, k, 1.0}; return true; }
      return false;
    };
    if (f == "id") { rc = needi("tag"); c = Col{0, iv["tag"].data(), 1, 0, 1.0}; }
    else if (f == "type") { rc = needi("type"); c = Col{0, iv["type"].data(), 1, 0, 1.0}; }
    else if (f == "x" || f == "y" || f == "z") { rc = needd("x", 3); c = Col{1, dv["x"].data(), 3, f == "x" ? 0 : f == "y" ? 1 : 2, 1.0}; }
    else if (vec3("v", "v") || vec3("f", "f") || vec3("omega", "omega") || vec3("torque", "tq")) {}
    else if (f == "radius") { rc = needd("radius", 1); c = Col{1, dv["radius"].data(), 1, 0, 1.0}; }
    else if (f == "diameter") { rc = needd("radius", 1); c = Col{1, dv["radius"].data(), 1, 0, 2.0}; }
    else if (f == "mass") { rc = needd("rmass", 1); c = Col{1, dv["rmass"].data(), 1, 0, 1.0}; }
    else if (f == "density") { rc = needd("density", 1); c = Col{1, dv["density"].data(), 1, 0, 1.0}; }
    else { fclose(fp); return fail(d, ERR_UNSUPPORTED, "dump custom field '%s' is outside the hot-path scope", f.c_str()); }
    if (rc) { fclose(fp); return engine_failed(d, rc); }
    cols.push_back(c);
  }
  for (long i = 0; i < n; i++) {
    for (auto &c : cols) {
      if (c.kind == 0) fprintf(fp, "%d ", ((const int *)c.p)[i]);
      else fprintf(fp, "%g ", c.scale * ((const double *)c.p)[(size_t)i * c.w + c.c]);
    }
    fprintf(fp, "\n");
  }
  fclose(fp);
  D.last = d->ntimestep;
  return OK;
}
int write_dumps_due(Deck *d, bool run_start = false)
{  // snapshots on the steps that are multiples of N, and at the start of a run when `dump_modify first yes` (output.cpp:150-190), once per step
  for (auto &kv : d->dumps) {
    Deck::Dump &D = kv.second;
    if (D.every > 0 && (d->ntimestep % D.every == 0 || (run_start && D.first)) && D.last != d->ntimestep) { const int rc = write_dump(d, D); if (rc) return rc; }
  }
  return OK;
}


// ---- particle insertion (host side): regions, templates, distributions, fix insert/pack --------------------------------
// The reference draws positions on the host from Park-Miller streams; what reaches the particle hot path is a list of new
// spheres inside one timestep.  This block restates the drawing so that a deck seeded like a reference deck creates the same
// spheres: region.cpp:498-700 (random points, Monte-Carlo volume), fix_insert_pack.cpp:337-597 (how many, where),
// fix_particledistribution_discrete.cpp:383-455 (which template), fix_insert.cpp:672-905 (the insertion step).
bool is_prime(int n) { if (n < 2) return false; for (long q = 2; q * q <= n; q++) if (n % q == 0) return false; return true; }
// Random::Random (random.cpp:55-88) + Input::add_and_validate_seed (input.cpp:1920-1953): a seed that is not a prime > 10000
// would be replaced by one of the reference's built-in seeds -- refused here instead of silently drawing another stream
int take_seed(Deck *d, const std::string &w, int &seed)
{
  int rc = inumeric(d, w, seed); if (rc) return rc;
  if (atol(w.c_str()) != (long)seed) return fail(d, ERR_ARG, "Seed %s is larger than INT_MAX", w.c_str());
  if (seed < 10000 || !is_prime(seed)) return fail(d, ERR_UNSUPPORTED, "seed %d: LIGGGHTS requires seeds to be prime numbers > 10000 (the reference would substitute one of its built-in seeds)", seed);
  return OK;
}
bool region_inside(const Deck::Region &R, double x, double y, double z)
{
  if (R.kind == 0) return x >= R.lo[0] && x <= R.hi[0] && y >= R.lo[1] && y <= R.hi[1] && z >= R.lo[2] && z <= R.hi[2];  // region_block.cpp:272-277
  double del1, del2, ax;  // region_cylinder.cpp:213-239
  if (R.axis == 'x') { del1 = y - R.c1; del2 = z - R.c2; ax = x; } else if (R.axis == 'y') { del1 = x - R.c1; del2 = z - R.c2; ax = y; } else { del1 = x - R.c1; del2 = y - R.c2; ax = z; }
  const double dist = std::sqrt(del1 * del1 + del2 * del2);
  return dist <= R.rad && ax >= R.clo && ax <= R.chi;
}
// Region::match_cut for an interior region: the point lies within `cut` of the surface (surface_interior() > 0,
// region_block.cpp:286-345, region_cylinder.cpp:249-359)
bool region_near_surface(const Deck::Region &R, const double *x, double cut)
{
  if (R.kind == 0) {
    if (x[0] < R.lo[0] || x[0] > R.hi[0] || x[1] < R.lo[1] || x[1] > R.hi[1] || x[2] < R.lo[2] || x[2] > R.hi[2]) return false;
    for (int k = 0; k < 3; k++) if (x[k] - R.lo[k] < cut || R.hi[k] - x[k] < cut) return true;
    return false;
  }
  double del1, del2, ax;
  if (R.axis == 'x') { del1 = x[1] - R.c1; del2 = x[2] - R.c2; ax = x[0]; } else if (R.axis == 'y') { del1 = x[0] - R.c1; del2 = x[2] - R.c2; ax = x[1]; } else { del1 = x[0] - R.c1; del2 = x[1] - R.c2; ax = x[2]; }
  const double r = std::sqrt(del1 * del1 + del2 * del2);
  if (r > R.rad || ax < R.clo || ax > R.chi) return false;
  if (R.rad - r < cut && r > 0.0) return true;
  return ax - R.clo < cut || R.chi - ax < cut;
}
bool in_domain(const Deck *d, const double *p) { for (int k = 0; k < 3; k++) if (!(p[k] >= d->lo[k] && p[k] <= d->hi[k])) return false; return true; }  // domain_I.h:51-63
bool in_subdomain(const Deck *d, const double *p) { for (int k = 0; k < 3; k++) if (!(p[k] >= d->lo[k] - 1.0e-8 && p[k] < d->hi[k] + 1.0e-8)) return false; return true; }  // domain_I.h:65-87, one process
// Region::volume_mc (region.cpp:633-700), one process
int region_volume_mc(Deck *d, Deck::Region &R, int n_test, bool cutflag, double cut, double &vol_global, double &vol_local)
{
  long n_in_local = 0, n_in_global = 0;
  for (int i = 0; i < n_test; i++) {
    double pos[3];
    for (int k = 0; k < 3; k++) pos[k] = R.ext_lo[k] + R.random.uniform() * (R.ext_hi[k] - R.ext_lo[k]);
    if (!in_domain(d, pos)) continue;
    if (region_inside(R, pos[0], pos[1], pos[2])) {
      n_in_global++;
      if (in_subdomain(d, pos) && !(cutflag && region_near_surface(R, pos, cut))) n_in_local++;
    }
  }
  const char *msg = "Unable to calculate region volume. Possible sources of error: (a) region volume is too small or out of domain, (b) particles for insertion are too large when using all_in yes, (c) region is 2d, but should be 3d";
  if (n_in_global == 0) return fail(d, ERR_ARG, "%s", msg);
  const double vol_bbox = (R.ext_hi[0] - R.ext_lo[0]) * (R.ext_hi[1] - R.ext_lo[1]) * (R.ext_hi[2] - R.ext_lo[2]);
  vol_global = static_cast<double>(n_in_global) / static_cast<double>(n_test * 1) * vol_bbox;
  vol_local = static_cast<double>(n_in_local) / static_cast<double>(n_test) * vol_bbox;
  const double vol_local_all = vol_local;
  if (vol_local_all < 1.e-10) return fail(d, ERR_ARG, "%s", msg);  // Region::volume_limit_
  vol_local *= (vol_global / vol_local_all);
  return OK;
}
// Region::generate_random / generate_random_shrinkby_cut with subdomain_flag (region.cpp:505-572)
int region_random_point(Deck *d, Deck::Region &R, bool shrink, double cut, double *pos)
{
  double lo[3], hi[3], diff[3];
  for (int k = 0; k < 3; k++) { lo[k] = std::max(R.ext_lo[k], d->lo[k]); hi[k] = std::min(R.ext_hi[k], d->hi[k]); diff[k] = hi[k] - lo[k]; }
  if (lo[0] >= hi[0] || lo[1] >= hi[1] || lo[2] >= hi[2]) return fail(d, ERR_ARG, "Impossible to generate random points on wrong sub-domain");
  if (shrink) for (int k = 0; k < 3; k++) if (R.ext_hi[k] - R.ext_lo[k] < 2. * cut)
    return fail(d, ERR_ARG, "Impossible to generate random points within region - region too small (smaller than twice the particle cutoff)");
  do {
    pos[0] = lo[0] + R.random.uniform() * diff[0];
    pos[1] = lo[1] + R.random.uniform() * diff[1];
    pos[2] = lo[2] + R.random.uniform() * diff[2];
  } while (!region_inside(R, pos[0], pos[1], pos[2]) || (shrink && region_near_surface(R, pos, cut)));
  return OK;
}
// FixInsert::setup -> FixInsertPack::calc_insertion_properties (fix_insert.cpp:407-438, fix_insert_pack.cpp:284-327): once, in
// the setup of the first run after the fix was defined
int insert_setup(Deck *d, Deck::Insert &I)
{
  if (I.setup_done) return OK;
  I.setup_done = true;
  Deck::Region &R = d->regions[I.region];
  R.random.reset(I.seed + 12);  // SEED_OFFSET, Region::reset_random
  const Deck::Distribution &D = d->distributions[I.dist];
  int rc = region_volume_mc(d, R, I.ntry_mc, I.all_in != 0, D.maxrbound, I.region_volume, I.region_volume_local); if (rc) return rc;
  if (I.region_volume <= 0. || I.region_volume_local < 0. || (I.region_volume_local - I.region_volume) / I.region_volume > 1e-3)
    return fail(d, ERR_ARG, "Fix insert: Region volume calculation with MC failed");
  char buf[160]; snprintf(buf, sizeof buf, "INFO: Particle insertion %s: inserting every %ld steps\n", I.id.c_str(), I.insert_every);
  d->out += buf; if (d->screen) fputs(buf, stdout);
  return OK;
}
struct NewSpheres { std::vector<int> type, mask; std::vector<double> x, v, omega, radius, density; long n = 0; };
// one insertion of one fix insert/pack (FixInsert::pre_exchange, fix_insert.cpp:672-905) against the spheres in ex/er/em
// (positions after the first half step, radii, masses; spheres created earlier in this step included)
int insert_pass(Deck *d, Deck::Insert &I, std::vector<double> &ex, std::vector<double> &er, std::vector<double> &em, NewSpheres &out)
{
  Deck::Region &R = d->regions[I.region];
  Deck::Distribution &D = d->distributions[I.dist];
  const long step = d->ntimestep + 1;
  // FixInsertPack::calc_ninsert_this (fix_insert_pack.cpp:337-434)
  if (I.warn_region) {
    double mn[3], mx[3];
    for (int k = 0; k < 3; k++) { mn[k] = R.ext_lo[k] + 1e-8; mx[k] = R.ext_hi[k] - 1e-8; }
    if (!in_domain(d, mn) || !in_domain(d, mx)) for (int k = 0; k < 3; k++) if (d->bstr[k].find('f') != std::string::npos)
      return fail(d, ERR_ARG, "Insertion region extends outside simulation box and a fixed boundary is used.Please use non-fixed boundaries in this case only");
  }
  long np_region = 0; double vol_region = 0., mass_region = 0.;
  const double _4Pi3 = 4. * M_PI / 3.;
  for (size_t i = 0; i < er.size(); i++) if (region_inside(R, ex[3 * i], ex[3 * i + 1], ex[3 * i + 2])) { np_region++; vol_region += _4Pi3 * er[i] * er[i] * er[i]; mass_region += em[i]; }
  int ninsert_this = 0;
  if (I.volumefraction > 0.) {
    ninsert_this = static_cast<int>((I.volumefraction * I.region_volume - vol_region) / D.volexpect + I.random.uniform());
    I.insertion_ratio = vol_region / (I.volumefraction * I.region_volume);
  } else if (I.ntotal > 0) {
    ninsert_this = (int)(I.ntotal - np_region);
    I.insertion_ratio = static_cast<double>(np_region) / static_cast<double>(I.ntotal);
  } else {
    ninsert_this = static_cast<int>((I.masstotal - mass_region) / D.massexpect + I.random.uniform());
    I.insertion_ratio = mass_region / I.masstotal;
  }
  if (ninsert_this < -200000) return fail(d, ERR_ARG, "overflow in particle number calculation: inserting too many particles in one step");
  if (ninsert_this < 0) ninsert_this = 0;
  if (I.insertion_ratio < 0.) I.insertion_ratio = 0.;
  if (I.insertion_ratio > 1.) I.insertion_ratio = 1.;
  // FixInsertPack::insertion_fraction (fix_insert_pack.cpp:438-445): a shrink-wrapped box (boundary m / s) counts as changing,
  // the Monte-Carlo volume is drawn again before every insertion
  bool box_change = false;
  for (int k = 0; k < 3; k++) if (d->bstr[k].find('m') != std::string::npos || d->bstr[k].find('s') != std::string::npos) box_change = true;
  if (box_change) { const int rc = region_volume_mc(d, R, I.ntry_mc, I.all_in != 0, D.maxrbound, I.region_volume, I.region_volume_local); if (rc) return rc; }
  // distribute_ninsert_this (fix_insert.cpp:912-987) with one process: everything is mine unless my share of the region is < 2 %
  if (I.exact_number) { if (I.region_volume_local / I.region_volume < 0.02) return fail(d, ERR_ARG, "Internal error distributing particles to processes"); }
  else ninsert_this = static_cast<int>(I.region_volume_local / I.region_volume * static_cast<double>(ninsert_this) + I.random.uniform());
  // FixParticledistributionDiscrete::randomize_list (fix_particledistribution_discrete.cpp:383-455)
  const int nt = (int)D.templates.size();
  std::vector<int> parttogen(nt);
  if (!I.exact_number) for (int i = 0; i < nt; i++) parttogen[i] = static_cast<int>(static_cast<double>(ninsert_this) * D.weight[i] + D.random.uniform());
  else {
    int truncated = 0; std::vector<double> remainder(nt);
    for (int i = 0; i < nt; i++) {
      parttogen[i] = static_cast<int>(static_cast<double>(ninsert_this) * D.weight[i]);
      truncated += parttogen[i];
      remainder[i] = static_cast<double>(ninsert_this) * D.weight[i] - static_cast<double>(parttogen[i]);
    }
    const int gap = ninsert_this - truncated;
    for (int i = 0; i < gap; i++) {
      const double r = D.random.uniform() * static_cast<double>(gap);
      int j = 0; double rsum = remainder[0];
      while (rsum < r && j < nt - 1) { j++; rsum += remainder[j]; }
      parttogen[j]++;
    }
  }
  std::vector<int> list;  // template index of every sphere to insert, large templates first
  for (int i = 0; i < nt; i++) for (int j = 0; j < parttogen[D.order[i]]; j++) list.push_back(D.order[i]);
  const int nlist = (int)list.size();
  auto schedule = [&](bool some) {
    if (some) { if (I.insert_every) I.next += I.insert_every; else I.next = 0; }
    else { if (I.insert_every) I.next += I.insert_every; else I.next = -1; }
  };
  if (nlist == 0) { schedule(false); return OK; }
  // FixInsertPack::x_v_omega (fix_insert_pack.cpp:474-597)
  const long maxtry = I.insertion_ratio >= 1. ? (long)nlist * I.maxattempt : (long)static_cast<int>(static_cast<double>(nlist * I.maxattempt) / (1. - I.insertion_ratio));
  long ntry = 0; int ninserted = 0; double mass_inserted = 0.;
  const size_t ex0 = er.size();
  auto velocity = [&](double *v) {  // FixInsert::generate_random_velocity (fix_insert.cpp:1023-1036)
    v[0] = I.v[0]; v[1] = I.v[1]; v[2] = I.v[2];
    if (I.vmode == 1) for (int k = 0; k < 3; k++) v[k] = I.v[k] + I.vfluct[k] * 2.0 * (I.random.uniform() - 0.50);
    else if (I.vmode == 2) for (int k = 0; k < 3; k++) v[k] = I.v[k] + I.vfluct[k] * I.random.gaussian();
  };
  // overlap search: uniform bins over the region's box, cell >= the largest diameter in play (RegionNeighborList::hasOverlap,
  // region_neighbor_list_I.h:72-98, tests rsq <= radsum^2 against every sphere that can reach)
  double rmaxall = D.maxrbound; for (double r : er) rmaxall = std::max(rmaxall, r);
  const double cell = 2.0 * rmaxall * 1.0001;
  double blo[3], cs[3]; int nb[3];
  for (int k = 0; k < 3; k++) {
    const double span = R.ext_hi[k] - R.ext_lo[k] + 4.0 * cell;
    cs[k] = std::max(cell, span / 200.0); blo[k] = R.ext_lo[k] - 2.0 * cell; nb[k] = (int)(span / cs[k]) + 1;
  }
  std::vector<std::vector<int>> bins((size_t)nb[0] * nb[1] * nb[2]);
  auto binof = [&](const double *p, int *b) -> bool { for (int k = 0; k < 3; k++) { const double q = (p[k] - blo[k]) / cs[k]; if (!(q >= 0.0) || q >= nb[k]) return false; b[k] = (int)q; } return true; };
  if (I.check_ol) for (size_t i = 0; i < er.size(); i++) { int b[3]; if (binof(&ex[3 * i], b)) bins[((size_t)b[2] * nb[1] + b[1]) * nb[0] + b[0]].push_back((int)i); }
  auto overlaps = [&](const double *p, double r) {
    int b[3]; if (!binof(p, b)) return false;
    for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
      const int bx = b[0] + dx, by = b[1] + dy, bz = b[2] + dz;
      if (bx < 0 || by < 0 || bz < 0 || bx >= nb[0] || by >= nb[1] || bz >= nb[2]) continue;
      for (int j : bins[((size_t)bz * nb[1] + by) * nb[0] + bx]) {
        const double del[3] = {p[0] - ex[3 * j], p[1] - ex[3 * j + 1], p[2] - ex[3 * j + 2]};
        const double rsq = del[0] * del[0] + del[1] * del[1] + del[2] * del[2], radsum = r + er[j];
        if (rsq <= radsum * radsum) return true;
      }
    }
    return false;
  };
  auto accept = [&](int t, const double *pos, const double *v) {
    const Deck::Template &T = d->templates[D.templates[t]];
    const double volume = T.radius * T.radius * T.radius * 4. * M_PI / 3., mass = T.density * volume;  // fix_template_sphere.cpp:349-350
    out.type.push_back(T.atom_type); out.mask.push_back(1 | T.groupbit | D.groupbit | I.groupbit);
    for (int k = 0; k < 3; k++) { out.x.push_back(pos[k]); out.v.push_back(v[k]); out.omega.push_back(I.omega[k]); ex.push_back(pos[k]); }
    out.radius.push_back(T.radius); out.density.push_back(T.density); out.n++;
    er.push_back(T.radius); em.push_back(mass);
    int b[3]; if (I.check_ol && binof(pos, b)) bins[((size_t)b[2] * nb[1] + b[1]) * nb[0] + b[0]].push_back((int)er.size() - 1);
    mass_inserted += mass; ninserted++;
  };
  double pos[3], v[3];
  if (!I.check_ol) {
    for (int it = 0; it < nlist; it++) {
      const double rbound = d->templates[D.templates[list[ninserted]]].radius;
      int rc = region_random_point(d, R, I.all_in != 0, rbound, pos); if (rc) return rc;
      velocity(v);
      if (pos[0] == 0. && pos[1] == 0. && pos[2] == 0.) return fail(d, ERR_ARG, "FixInsertPack::x_v_omega() illegal position");
      accept(list[ninserted], pos, v);
    }
  } else {
    while (ntry < maxtry && ninserted < nlist) {
      const int t = list[ninserted];
      const double rbound = d->templates[D.templates[t]].radius;
      bool placed = false;
      while (!placed && ntry < maxtry) {
        bool clear;
        do {
          int rc = region_random_point(d, R, I.all_in != 0, rbound, pos); if (rc) return rc;
          ntry++;
          clear = true;  // Domain::check_dist_subbox_borders (domain_I.h:127-149): keep rbound away from the sub-box faces
          for (int k = 0; k < 3; k++) if (std::fabs(d->lo[k] - pos[k]) < rbound || std::fabs(d->hi[k] - pos[k]) < rbound) clear = false;
        } while (I.check_border && ntry < maxtry && !clear);
        if (ntry == maxtry) break;
        velocity(v);
        if (!overlaps(pos, rbound)) { accept(t, pos, v); placed = true; }
      }
    }
  }
  (void)ex0;
  I.ninserted += ninserted; I.massinserted += mass_inserted;
  char buf[320];
  snprintf(buf, sizeof buf, "INFO: Particle insertion %s: inserted %d particle templates (mass %e) at step %ld\n - a total of %ld particle templates (mass %e) inserted so far.\n",
           I.id.c_str(), ninserted, mass_inserted, step, I.ninserted, I.massinserted);
  d->out += buf; if (d->screen) fputs(buf, stdout);
  if (ninserted < nlist) d->warnings += "Particle insertion: Less insertions than requested\n";
  schedule(true);
  return OK;
}
// the timestep d->ntimestep + 1 with every fix insert/pack that is due on it
int insertion_step(Deck *d)
{
  const long step = d->ntimestep + 1;
  std::vector<double> ex, er, em;
  TRY(API(insert_step_begin)(d->e));
  const long n0 = d->uploaded ? API(nlocal)(d->e) : 0;
  int maxtag = 0;
  if (n0) {
    if (d->nprocs > 1) return fail(d, ERR_UNSUPPORTED, "fix insert/pack into a box that already holds particles needs all of them on one process: several bricks are outside the hot-path scope");
    ex.resize(3 * n0); er.resize(n0); em.resize(n0);
    std::vector<int> tg(n0);
    TRY(API(download)(d->e, "x", ex.data(), n0)); TRY(API(download)(d->e, "radius", er.data(), n0)); TRY(API(download)(d->e, "rmass", em.data(), n0));
    TRY(API(download)(d->e, "tag", tg.data(), n0));
    for (int t : tg) maxtag = std::max(maxtag, t);  // Atom::tag_extend continues behind the largest id present
  }
  NewSpheres S;
  for (auto &I : d->inserts) if (I.next == step) { const int rc = insert_pass(d, I, ex, er, em, S); if (rc) return rc; }
  std::vector<int> tag(S.n);
  for (long i = 0; i < S.n; i++) tag[i] = maxtag + 1 + (int)i;
  if (S.n && !d->newton_off && d->have_pair) return fail(d, ERR_ARG, "Pair granular with shear history requires newton pair off");
  TRY(API(insert_step_end)(d->e, S.n, tag.data(), S.type.data(), S.mask.data(), S.x.data(), S.v.data(), S.omega.data(), S.radius.data(), S.density.data()));
  if (S.n) { d->uploaded = true; d->maxtag = std::max(d->maxtag, maxtag + (int)S.n); }
  return OK;
}

int first_run_prepare(Deck *d)
{
  if (d->uploaded) {
    if (!d->ptag.empty()) {  // atoms created between two runs
      TRY(API(insert_particles)(d->e, (long)d->ptag.size(), d->ptag.data(), d->ptype.data(), d->pmask.data(), d->px.data(), d->pv.data(), d->pomega.data(),
                                d->pradius.data(), d->pdensity.data()));
      d->ptag.clear(); d->ptype.clear(); d->pmask.clear(); d->px.clear(); d->pv.clear(); d->pomega.clear(); d->pradius.clear(); d->pdensity.clear();
    }
    return OK;
  }
  if (!d->have_box) return fail(d, ERR_STATE, "Run command before simulation box is defined");
  if (d->tag.empty()) {
    if (!d->inserts.empty()) return OK;  // an empty box that fix insert/pack fills
    return fail(d, ERR_UNSUPPORTED, "no particles: initial states come from read_data, create_atoms single or fix insert/pack");
  }
  if (!d->newton_off && d->have_pair) return fail(d, ERR_ARG, "Pair granular with shear history requires newton pair off");
  TRY(API(upload_particles)(d->e, (long)d->tag.size(), d->tag.data(), d->type.data(), d->mask.data(), d->x.data(), d->v.data(), d->omega.data(),
                            d->radius.data(), d->density.data()));
  d->uploaded = true;
  return OK;
}
int send_box(Deck *d)
{
  if (d->box_sent || !d->have_box) return OK;
  TRY(API(set_box)(d->e, d->lo, d->hi, d->periodic));
  TRY(API(set_ntypes)(d->e, d->ntypes));
  d->box_sent = true;
  return OK;
}

int cmd_fix(Deck *d, const std::vector<std::string> &w)
{
  if (w.size() < 4) return fail(d, ERR_ARG, "Illegal fix command");
  const std::string &id = w[1], &group = w[2], &style = w[3];
  if (!d->groups.count(group)) return fail(d, ERR_ARG, "Could not find fix group ID %s", group.c_str());
  const int bit = d->groups[group];
  int rc = send_box(d); if (rc) return rc;
  // the reference applies its post_force fixes in the order of their definition; here fix freeze always acts last
  if (d->have_freeze && (style == "gravity" || style == "addforce" || style == "viscous" || style == "wall/gran"))
    d->warnings += "fix " + style + " defined after fix freeze: the reference would let it act on the frozen group, here frozen particles stay force free\n";
  if (style == "property/global") {  // fix_property_global.cpp:60-200
    if (w.size() < 7) return fail(d, ERR_ARG, "Illegal fix property/global command, not enough arguments");
    const std::string &name = w[4], &kind = w[5];
    size_t from = 6;
    if (kind == "peratomtypepair" || kind == "atomtypepair") {
      int nt; rc = inumeric(d, w[6], nt); if (rc) return rc;
      if (nt != d->ntypes) return fail(d, ERR_ARG, "Fix property/global %s: the number of atom types (%d) must match create_box / the data file (%d)", name.c_str(), nt, d->ntypes);
      from = 7;
    }
    std::vector<double> vals;
    for (size_t k = from; k < w.size(); k++) { double v; rc = numeric(d, w[k], v); if (rc) return rc; vals.push_back(v); }
    const char *kk = (kind == "atomtypepair") ? "peratomtypepair" : (kind == "atomtype") ? "peratomtype" : kind.c_str();
    TRY(API(set_property)(d->e, name.c_str(), kk, vals.data(), (int)vals.size()));
    return OK;
  }
  if (style == "gravity") {  // fix_gravity.cpp:60-140
    if (w.size() != 9 || w[5] != "vector") return fail(d, w.size() >= 6 && w[5] != "vector" ? ERR_UNSUPPORTED : ERR_ARG, "Illegal fix gravity command (supported: MAG vector x y z)");
    if (bit != 1) return fail(d, ERR_UNSUPPORTED, "fix gravity on a group other than 'all' is outside the hot-path scope");
    double mag, dir[3];
    rc = numeric(d, w[4], mag); if (rc) return rc;
    for (int k = 0; k < 3; k++) { rc = numeric(d, w[6 + k], dir[k]); if (rc) return rc; }
    TRY(API(set_gravity)(d->e, mag, dir));
    return OK;
  }
  if (style == "wall/gran") {  // fix_wall_gran.cpp:120-342
    if (bit != 1) return fail(d, ERR_UNSUPPORTED, "fix wall/gran on a group other than 'all' is outside the hot-path scope");
    bool mesh = false;
    for (size_t k = 4; k < w.size(); k++) if (w[k] == "mesh") mesh = true; else if (w[k] == "primitive") break;
    std::vector<const char *> a = cptrs(w, 4);
    if (mesh) TRY(API(add_wall_mesh)(d->e, id.c_str(), (int)a.size(), a.data()));
    else TRY(API(add_wall_primitive)(d->e, id.c_str(), (int)a.size(), a.data()));
    return OK;
  }
  if (style.compare(0, 12, "mesh/surface") == 0) {  // any mesh/surface* style falls back to the base creator (modify.cpp:852-856)
    if (w.size() < 6 || w[4] != "file") return fail(d, w.size() >= 5 && w[4] == "fix" ? ERR_UNSUPPORTED : ERR_ARG, "expecting keyword 'file' or 'fix'");
    std::vector<double> nodes;
    rc = read_stl(d, w[5], nodes); if (rc) return rc;
    int atom_type = 1;
    std::vector<std::string> pass;
    for (size_t k = 6; k < w.size();) {
      const std::string &key = w[k];
      auto need = [&](size_t n) { return k + n < w.size() + 0 ? OK : fail(d, ERR_ARG, "not enough arguments for '%s'", key.c_str()); };
      if (key == "type") { rc = need(1); if (rc) return rc; rc = inumeric(d, w[k + 1], atom_type); if (rc) return rc; if (atom_type < 1) return fail(d, ERR_ARG, "'type' > 0 required"); k += 2; }
      else if (key == "move") {
        rc = need(3); if (rc) return rc;
        double dx[3]; for (int c = 0; c < 3; c++) { rc = numeric(d, w[k + 1 + c], dx[c]); if (rc) return rc; }
        for (size_t n = 0; n < nodes.size(); n++) nodes[n] = nodes[n] + dx[n % 3];
        k += 4;
      } else if (key == "scale") {
        rc = need(1); if (rc) return rc;
        double s; rc = numeric(d, w[k + 1], s); if (rc) return rc;
        for (size_t n = 0; n < nodes.size(); n++) nodes[n] *= s;
        k += 2;
      } else if (key == "rotate") {
        rc = need(6); if (rc) return rc;
        if (w[k + 1] != "axis") return fail(d, ERR_ARG, "expecting keyword 'axis' after keyword 'rotate'");
        if (w[k + 5] != "angle") return fail(d, ERR_ARG, "expecting keyword 'angle' after axis definition");
        double ax[3], ang; for (int c = 0; c < 3; c++) { rc = numeric(d, w[k + 2 + c], ax[c]); if (rc) return rc; }
        rc = numeric(d, w[k + 6], ang); if (rc) return rc;
        if (std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]) < 1e-5) return fail(d, ERR_ARG, "illegal magnitude of rotation axis");
        mesh_rotate(nodes, ax, ang);
        k += 7;
      } else if (key == "curvature" || key == "precision" || key == "stress") { rc = need(1); if (rc) return rc; pass.push_back(key); pass.push_back(w[k + 1]); k += 2; }
      else if (key == "reference_point") { rc = need(3); if (rc) return rc; pass.push_back(key); for (int c = 1; c <= 3; c++) pass.push_back(w[k + c]); k += 4; }
      else if (key == "verbose" || key == "heal") { rc = need(1); if (rc) return rc; k += 2; }
      else return fail(d, ERR_UNSUPPORTED, "fix mesh/surface keyword '%s' is outside the hot-path scope", key.c_str());
    }
    if (style == "mesh/surface/stress") {  // the stress module tracks the total force unless the deck says `stress off`
      bool given = false;
      for (size_t k = 0; k + 1 < pass.size(); k++) if (pass[k] == "stress") given = true;
      if (!given) { pass.push_back("stress"); pass.push_back("on"); }
    }
    std::vector<const char *> a = cptrs(pass, 0);
    TRY(API(add_mesh)(d->e, id.c_str(), atom_type, nodes.data(), (long)(nodes.size() / 9), (int)a.size(), a.data()));
    return OK;
  }
  if (style == "move/mesh") {  // fix_move_mesh.cpp:60-110
    if (w.size() < 7 || w[4] != "mesh") return fail(d, ERR_ARG, "Illegal fix move/mesh command, expecting keyword 'mesh'");
    std::vector<const char *> a = cptrs(w, 6);
    TRY(API(move_mesh)(d->e, w[5].c_str(), (int)a.size(), a.data()));
    return OK;
  }
  if (style == "freeze") { TRY(API(set_freeze)(d->e, bit)); d->have_freeze = true; return OK; }
  if (style == "addforce") {  // fix_addforce.cpp:60-130 (constant components; variables / region / energy options are not on the path)
    if (w.size() != 7) return fail(d, w.size() < 7 ? ERR_ARG : ERR_UNSUPPORTED, w.size() < 7 ? "Illegal fix addforce command" : "fix addforce options are outside the hot-path scope");
    double f[3]; for (int k = 0; k < 3; k++) { if (w[4 + k].compare(0, 2, "v_") == 0) return fail(d, ERR_UNSUPPORTED, "fix addforce with a variable is outside the hot-path scope"); rc = numeric(d, w[4 + k], f[k]); if (rc) return rc; }
    TRY(API(set_extra_force)(d->e, id.c_str(), 0, bit, f, 3)); d->extra_fixes.insert(id); return OK;
  }
  if (style == "viscous") {  // fix_viscous.cpp:40-95 (one gamma; per-type `scale` is not on the path)
    if (w.size() != 5) return fail(d, w.size() < 5 ? ERR_ARG : ERR_UNSUPPORTED, w.size() < 5 ? "Illegal fix viscous command" : "fix viscous scale is outside the hot-path scope");
    double g; rc = numeric(d, w[4], g); if (rc) return rc;
    TRY(API(set_extra_force)(d->e, id.c_str(), 1, bit, &g, 1)); d->extra_fixes.insert(id); return OK;
  }
  if (style == "nve/sphere") { TRY(API(set_integrate)(d->e, bit)); return OK; }
  if (style == "check/timestep/gran" || style == "print" || style.compare(0, 4, "ave/") == 0) {
    d->ignored_fixes[id] = style; d->warnings += "fix " + style + " ignored (diagnostic / output only)\n"; return OK;
  }
  if (style == "balance") {  // dynamic load balancing (fix_balance.cpp) moves brick boundaries, never the physics: bricks stay static here
    d->ignored_fixes[id] = style; d->warnings += "fix balance ignored (bricks are static)\n"; return OK;
  }
  if (style == "particletemplate/sphere") {  // fix_template_sphere.cpp:72-260 (radius / density: `constant` is the only style the reference has)
    if (w.size() < 5) return fail(d, ERR_ARG, "Illegal fix particletemplate/sphere command, not enough arguments");
    Deck::Template T; T.groupbit = bit;
    int seed; rc = take_seed(d, w[4], seed); if (rc) return rc;
    bool have_r = false, have_rho = false;
    for (size_t k = 5; k < w.size();) {
      if (w[k] == "atom_type" && k + 1 < w.size()) { rc = inumeric(d, w[k + 1], T.atom_type); if (rc) return rc; if (T.atom_type < 1) return fail(d, ERR_ARG, "invalid atom type (must be >=1)"); k += 2; }
      else if ((w[k] == "radius" || w[k] == "density") && k + 2 < w.size()) {
        if (w[k + 1] != "constant") return fail(d, ERR_ARG, "invalid %s random style", w[k].c_str());
        double val; rc = numeric(d, w[k + 2], val); if (rc) return rc;
        if (val <= 0.) return fail(d, ERR_ARG, "Illegal fix particletemplate/sphere command, %s must be >= 0", w[k].c_str());
        if (w[k] == "radius") { T.radius = val; have_r = true; } else { T.density = val; have_rho = true; }
        k += 3;
      } else if (w[k] == "volume_limit" && k + 1 < w.size()) k += 2;
      else return fail(d, ERR_UNSUPPORTED, "fix particletemplate/sphere keyword '%s' is outside the hot-path scope", w[k].c_str());
    }
    if (!have_rho) return fail(d, ERR_ARG, "have to define 'density'");
    if (!have_r) return fail(d, ERR_ARG, "have to define 'radius'");
    if (T.atom_type > d->ntypes) return fail(d, ERR_ARG, "Invalid atom type in fix particletemplate/sphere");
    T.volexpect = T.radius * T.radius * T.radius * 4. * M_PI / 3.;  // cubic_expectancy * 4 pi / 3 (fix_template_sphere.cpp:256-257)
    T.massexpect = T.density * T.volexpect;
    d->templates[id] = T; return OK;
  }
  if (style.compare(0, 29, "particledistribution/discrete") == 0) {  // fix_particledistribution_discrete.cpp:75-245
    if (w.size() < 8) return fail(d, ERR_ARG, "Illegal fix particledistribution/discrete command, not enough arguments");
    Deck::Distribution D; D.groupbit = bit;
    const bool mass_based = style.find("numberbased") == std::string::npos;
    int seed; rc = take_seed(d, w[4], seed); if (rc) return rc;
    D.random.reset(seed);
    int nt; rc = inumeric(d, w[5], nt); if (rc) return rc;
    if (nt < 1) return fail(d, ERR_ARG, "illegal number of templates");
    if ((int)w.size() != 6 + 2 * nt) return fail(d, ERR_ARG, "# of templates does not match # of arguments");
    for (int i = 0; i < nt; i++) {
      const std::string &tid = w[6 + 2 * i];
      if (!d->templates.count(tid)) return fail(d, ERR_ARG, "invalid ID for fix particletemplate provided");
      for (auto &o : D.templates) if (o == tid) return fail(d, ERR_ARG, "cannot use the same template twice");
      double wt; rc = numeric(d, w[7 + 2 * i], wt); if (rc) return rc;
      if (wt < 0) return fail(d, ERR_ARG, "invalid weight");
      D.templates.push_back(tid); D.weight.push_back(wt);
    }
    double sum = 0; for (double x : D.weight) sum += x;
    if (std::fabs(sum - 1.) > 0.00001) d->warnings += "particledistribution/discrete: sum of distribution weights != 1, normalizing distribution\n";
    for (double &x : D.weight) x /= sum;
    if (mass_based) {  // mass-% to number-%
      for (int i = 0; i < nt; i++) D.weight[i] = D.weight[i] / d->templates[D.templates[i]].massexpect;
      sum = 0; for (double x : D.weight) sum += x;
      for (double &x : D.weight) x /= sum;
    }
    for (int i = 0; i < nt; i++) {
      const Deck::Template &T = d->templates[D.templates[i]];
      D.volexpect += T.volexpect * D.weight[i]; D.massexpect += T.massexpect * D.weight[i];
      D.maxrbound = std::max(D.maxrbound, T.radius);
    }
    D.order.resize(nt); for (int i = 0; i < nt; i++) D.order[i] = i;  // bubble sort by insertion volume, descending (:213-232)
    bool swapped; int n = nt;
    do {
      swapped = false;
      for (int i = 0; i < nt - 1; i++)
        if (d->templates[D.templates[D.order[i]]].volexpect < d->templates[D.templates[D.order[i + 1]]].volexpect) { std::swap(D.order[i], D.order[i + 1]); swapped = true; }
      n--;
    } while (swapped && n > 0);
    d->distributions[id] = D; return OK;
  }
  if (style == "insert/pack") {  // fix_insert.cpp:79-395, fix_insert_pack.cpp:74-200
    if (w.size() < 7) return fail(d, ERR_ARG, "not enough arguments");
    if (w[4] != "seed") return fail(d, ERR_ARG, "expecting keyword 'seed'");
    Deck::Insert I; I.id = id; I.groupbit = bit;
    rc = take_seed(d, w[5], I.seed); if (rc) return rc;
    I.random.reset(I.seed);
    I.next = d->ntimestep + 1;  // first insertion on the next timestep
    auto yesno = [&](const std::string &v, int &out) -> int { if (v == "yes") out = 1; else if (v == "no") out = 0; else return fail(d, ERR_ARG, "expecting 'yes' or 'no'"); return OK; };
    for (size_t k = 6; k < w.size();) {
      const std::string &kw = w[k];
      auto need = [&](size_t n) { return k + n < w.size(); };
      if (kw == "distributiontemplate" && need(1)) {
        if (!d->distributions.count(w[k + 1])) return fail(d, ERR_ARG, "Fix insert requires you to define a valid ID for a fix of type particledistribution/discrete");
        I.dist = w[k + 1]; k += 2;
      } else if (kw == "maxattempt" && need(1)) { rc = inumeric(d, w[k + 1], I.maxattempt); if (rc) return rc; k += 2; }
      else if ((kw == "insert_every" || kw == "every") && need(1)) {
        if (w[k + 1] == "once") I.insert_every = 0;
        else { int ie; rc = inumeric(d, w[k + 1], ie); if (rc) return rc; if (ie < 0) return fail(d, ERR_ARG, "insert_every must be >= 0"); I.insert_every = ie; }
        k += 2;
      } else if (kw == "start" && need(1)) {
        int st; rc = inumeric(d, w[k + 1], st); if (rc) return rc;
        if (st < d->ntimestep + 1) return fail(d, ERR_ARG, "'start' step can not be before current step");
        I.next = st; k += 2;
      } else if (kw == "overlapcheck" && need(1)) { rc = yesno(w[k + 1], I.check_ol); if (rc) return rc; k += 2; }
      else if (kw == "all_in" && need(1)) { rc = yesno(w[k + 1], I.all_in); if (rc) return rc; k += 2; }
      else if (kw == "random_distribute" && need(1)) { if (w[k + 1] == "uncorrelated") I.exact_number = 0; else if (w[k + 1] == "exact") I.exact_number = 1; else return fail(d, ERR_ARG, "Illegal fix insert command"); k += 2; }
      else if (kw == "verbose" && need(1)) k += 2;
      else if (kw == "vel" && need(4)) {
        const std::string &m = w[k + 1];
        if (m == "constant") { for (int q = 0; q < 3; q++) { rc = numeric(d, w[k + 2 + q], I.v[q]); if (rc) return rc; } k += 5; }
        else if ((m == "uniform" || m == "gaussian") && need(7)) {
          I.vmode = m == "uniform" ? 1 : 2;
          for (int q = 0; q < 3; q++) { rc = numeric(d, w[k + 2 + q], I.v[q]); if (rc) return rc; rc = numeric(d, w[k + 5 + q], I.vfluct[q]); if (rc) return rc; }
          k += 8;
        } else return fail(d, ERR_ARG, "expecting keyword 'constant' or 'uniform' or 'gaussian' after keyword 'vel'");
      } else if (kw == "omega" && need(4)) {
        if (w[k + 1] != "constant") return fail(d, ERR_ARG, "expecting keyword 'constant' after keyword 'omega'");
        for (int q = 0; q < 3; q++) { rc = numeric(d, w[k + 2 + q], I.omega[q]); if (rc) return rc; }
        k += 5;
      } else if (kw == "region" && need(1)) {
        if (d->opaque_regions.count(w[k + 1])) return fail(d, ERR_UNSUPPORTED, "fix insert/pack into a region that is neither a block nor a cylinder is outside the hot-path scope");
        if (!d->regions.count(w[k + 1])) return fail(d, ERR_ARG, "region ID does not exist");
        I.region = w[k + 1]; k += 2;
      } else if (kw == "volumefraction_region" && need(1)) { rc = numeric(d, w[k + 1], I.volumefraction); if (rc) return rc; if (I.volumefraction < 0. || I.volumefraction > 1.) return fail(d, ERR_ARG, "Invalid volumefraction"); k += 2; }
      else if (kw == "particles_in_region" && need(1)) { int nn; rc = inumeric(d, w[k + 1], nn); if (rc) return rc; if (nn <= 0) return fail(d, ERR_ARG, "'ntotal_region' > 0 required"); I.ntotal = nn; k += 2; }
      else if (kw == "mass_in_region" && need(1)) { rc = numeric(d, w[k + 1], I.masstotal); if (rc) return rc; if (I.masstotal <= 0.) return fail(d, ERR_ARG, "'masstotal_region' > 0 required"); k += 2; }
      else if (kw == "ntry_mc" && need(1)) { rc = inumeric(d, w[k + 1], I.ntry_mc); if (rc) return rc; if (I.ntry_mc < 1000) return fail(d, ERR_ARG, "ntry_mc must be > 1000"); k += 2; }
      else if (kw == "warn_region" && need(1)) { int v; rc = yesno(w[k + 1], v); if (rc) return rc; I.warn_region = v != 0; k += 2; }
      else if (kw == "check_dist_from_subdomain_border" && need(1)) { int v; rc = yesno(w[k + 1], v); if (rc) return rc; I.check_border = v != 0; k += 2; }
      else return fail(d, ERR_UNSUPPORTED, "fix insert/pack keyword '%s' is outside the hot-path scope", kw.c_str());
    }
    if (I.dist.empty()) return fail(d, ERR_ARG, "have to define a 'distributiontemplate'");
    if (I.region.empty()) return fail(d, ERR_ARG, "must define an insertion region");
    if (I.insert_every < 0) return fail(d, ERR_ARG, "must define 'insert_every'");
    if ((I.volumefraction > 0.) + (I.ntotal > 0) + (I.masstotal > 0.) != 1)
      return fail(d, ERR_ARG, "must define exactly one keyword out of 'volumefraction_region', 'particles_in_region', and 'mass_in_region'");
    d->inserts.push_back(I); return OK;
  }
  return fail(d, ERR_UNSUPPORTED, "fix style '%s' is outside the hot-path scope", style.c_str());
}

int cmd_group(Deck *d, const std::vector<std::string> &w)
{  // group.cpp:90-260 (styles id and type; a group keeps its bit, members are OR-ed in)
  if (w.size() < 4) return fail(d, ERR_ARG, "Illegal group command");
  if (d->uploaded) return fail(d, ERR_UNSUPPORTED, "group changes after the first run are outside the hot-path scope");
  int bit;
  if (d->groups.count(w[1])) bit = d->groups[w[1]];
  else { if (d->groups.size() >= 31) return fail(d, ERR_ARG, "Too many groups"); bit = 1 << (int)d->groups.size(); d->groups[w[1]] = bit; }
  if (w[2] == "region") {  // group.cpp:150-165: the atoms that are inside the region now
    if (d->opaque_regions.count(w[3])) return fail(d, ERR_UNSUPPORTED, "group region with a region that is neither a block nor a cylinder is outside the hot-path scope");
    if (!d->regions.count(w[3])) return fail(d, ERR_ARG, "Group region ID does not exist");
    const Deck::Region &R = d->regions[w[3]];
    for (size_t i = 0; i < d->tag.size(); i++) if (region_inside(R, d->x[3 * i], d->x[3 * i + 1], d->x[3 * i + 2])) d->mask[i] |= bit;
    return OK;
  }
  if (w[2] == "union" || w[2] == "subtract" || w[2] == "intersect") {  // group.cpp:330-460
    std::vector<int> bits;
    for (size_t k = 3; k < w.size(); k++) { if (!d->groups.count(w[k])) return fail(d, ERR_ARG, "Group ID does not exist"); bits.push_back(d->groups[w[k]]); }
    if (w[2] != "union" && bits.size() < 2) return fail(d, ERR_ARG, "Illegal group command");
    for (size_t i = 0; i < d->tag.size(); i++) {
      bool in;
      if (w[2] == "union") { in = false; for (int b : bits) in |= (d->mask[i] & b) != 0; }
      else if (w[2] == "subtract") { in = (d->mask[i] & bits[0]) != 0; for (size_t q = 1; q < bits.size(); q++) if (d->mask[i] & bits[q]) in = false; }
      else { in = true; for (int b : bits) if (!(d->mask[i] & b)) in = false; }
      if (in) d->mask[i] |= bit;
    }
    return OK;
  }
  if (w[2] != "id" && w[2] != "type") return fail(d, ERR_UNSUPPORTED, "group style '%s' is outside the hot-path scope", w[2].c_str());
  std::vector<std::pair<int, int>> ranges;
  for (size_t k = 3; k < w.size(); k++) {
    const size_t c = w[k].find(':');
    int a, b;
    int rc = inumeric(d, w[k].substr(0, c), a); if (rc) return rc;
    b = a;
    if (c != std::string::npos) { rc = inumeric(d, w[k].substr(c + 1), b); if (rc) return rc; }
    ranges.push_back({a, b});
  }
  const std::vector<int> &key = (w[2] == "id") ? d->tag : d->type;
  for (size_t i = 0; i < key.size(); i++) for (auto &r : ranges) if (key[i] >= r.first && key[i] <= r.second) { d->mask[i] |= bit; break; }
  return OK;
}

int run_file(Deck *d, const std::string &path);

int one(Deck *d, const std::string &raw)
{
  std::string line = raw;
  int rc = substitute(d, line); if (rc) return rc;
  const std::vector<std::string> w = split(line);
  if (w.empty()) return OK;
  const std::string &c = w[0];
  static const char *output_only[] = {"thermo_modify", "compute", "uncompute", "echo", "log",
                                     "print", "restart", "write_restart", "write_data", "info", "reset_timestep_info", "thermo_log", nullptr};
  for (int k = 0; output_only[k]; k++) if (c == output_only[k]) { d->warnings += c + " ignored (output only)\n"; return OK; }
  if (c == "variable") {  // styles equal (formula, see Formula) / string / index (variable.cpp:90-330)
    if (w.size() < 4) return fail(d, ERR_ARG, "Illegal variable command");
    if (w[2] != "equal" && w[2] != "string" && w[2] != "index") return fail(d, ERR_UNSUPPORTED, "variable style '%s' is outside the hot-path scope", w[2].c_str());
    if (w[2] == "index" && d->vars.count(w[1])) return OK;  // an index variable keeps its first value
    if (w[2] == "equal") {  // the formula is everything after the style (it may contain blanks); evaluated when substituted
      std::string f = w[3];
      for (size_t k = 4; k < w.size(); k++) f += w[k];
      d->vars[w[1]] = f; d->var_is_equal.insert(w[1]);
    } else { d->vars[w[1]] = w[3]; d->var_is_equal.erase(w[1]); }
    return OK;
  }
  if (c == "units") { if (w.size() != 2) return fail(d, ERR_ARG, "Illegal units command"); TRY(API(set_units)(d->e, w[1].c_str())); return OK; }
  if (c == "atom_style") {
    if (w.size() < 2) return fail(d, ERR_ARG, "Illegal atom_style command");
    if (w[1] != "sphere" && w[1] != "granular") return fail(d, ERR_UNSUPPORTED, "atom_style %s is outside the hot-path scope (sphere | granular)", w[1].c_str());
    return OK;
  }
  if (c == "atom_modify" || c == "pair_coeff" || c == "hard_particles" || c == "soft_particles" || c == "modify_timing") return OK;
  if (c == "dimension") { if (w.size() != 2 || w[1] != "3") return fail(d, ERR_UNSUPPORTED, "only dimension 3"); return OK; }
  if (c == "boundary") {
    if (w.size() != 4) return fail(d, ERR_ARG, "Illegal boundary command");
    for (int k = 0; k < 3; k++) {
      const std::string &b = w[1 + k];
      d->bstr[k] = b.size() == 1 ? b + b : b;
      if (b == "p") d->periodic[k] = 1;
      else if (b == "f" || b == "m" || b == "s" || b == "ff" || b == "mm" || b == "fm" || b == "mf" || b == "ss") d->periodic[k] = 0;
      else return fail(d, ERR_ARG, "Illegal boundary command");
    }
    return OK;
  }
  if (c == "newton") { if (w.size() < 2) return fail(d, ERR_ARG, "Illegal newton command"); d->newton_off = (w[1] == "off"); return OK; }
  if (c == "communicate" || c == "comm_modify") {
    for (size_t k = 1; k + 1 < w.size(); k++) if (w[k] == "vel" && w[k + 1] != "yes") return fail(d, ERR_ARG, "Pair granular requires ghost atoms store velocity");
    return OK;
  }
  if (c == "processors") {
    if (w.size() < 4) return fail(d, ERR_ARG, "Illegal processors command");
    if (w[1] == "*" || w[2] == "*" || w[3] == "*") return OK;  // left to the engine's own choice
    int p[3]; for (int k = 0; k < 3; k++) { rc = inumeric(d, w[1 + k], p[k]); if (rc) return rc; }
    TRY(API(set_processors)(d->e, p[0], p[1], p[2])); d->nprocs = p[0] * p[1] * p[2]; return OK;
  }
  if (c == "region") {
    Deck::Region R;
    R.random.reset(3012211);  // region.cpp:456-457
    size_t opt = 9;
    if (w.size() >= 3 && w[2] == "cylinder") {  // region_cylinder.cpp:64-190
      if (w.size() < 9) return fail(d, ERR_ARG, "Illegal region cylinder command");
      if (w[3] != "x" && w[3] != "y" && w[3] != "z") return fail(d, ERR_ARG, "Illegal region cylinder command");
      R.kind = 1; R.axis = w[3][0];
      rc = numeric(d, w[4], R.c1); if (rc) return rc; rc = numeric(d, w[5], R.c2); if (rc) return rc; rc = numeric(d, w[6], R.rad); if (rc) return rc;
      rc = numeric(d, w[7], R.clo); if (rc) return rc; rc = numeric(d, w[8], R.chi); if (rc) return rc;
      if (R.rad <= 0.0) return fail(d, ERR_ARG, "Illegal region cylinder command");
      const int a = R.axis == 'x' ? 0 : R.axis == 'y' ? 1 : 2, b1 = a == 0 ? 1 : 0, b2 = a == 2 ? 1 : 2;
      R.ext_lo[a] = R.clo; R.ext_hi[a] = R.chi; R.ext_lo[b1] = R.c1 - R.rad; R.ext_hi[b1] = R.c1 + R.rad; R.ext_lo[b2] = R.c2 - R.rad; R.ext_hi[b2] = R.c2 + R.rad;
      for (int k = 0; k < 3; k++) { R.lo[k] = R.ext_lo[k]; R.hi[k] = R.ext_hi[k]; }
    } else if (w.size() >= 3 && w[2] != "block") {  // other shapes only feed groups / insertion, which are rejected where they are used
      d->opaque_regions.insert(w[1]); d->warnings += "region " + w[1] + " (" + w[2] + ") kept as a name only\n"; return OK;
    } else {
      if (w.size() < 9) return fail(d, ERR_ARG, "Illegal region command");
      // INF = -/+ 1e20, EDGE = the box bound (region_block.cpp:70-230)
      auto bound = [&](const std::string &t, int k, bool hi, double &out) -> int {
        if (t == "INF" || t == "EDGE") {
          if (!d->have_box) return fail(d, ERR_ARG, "Cannot use region INF or EDGE when box does not exist");
          out = t == "INF" ? (hi ? 1.0e20 : -1.0e20) : (hi ? d->hi[k] : d->lo[k]); return OK;
        }
        return numeric(d, t, out);
      };
      for (int k = 0; k < 3; k++) { rc = bound(w[3 + 2 * k], k, false, R.lo[k]); if (rc) return rc; rc = bound(w[4 + 2 * k], k, true, R.hi[k]); if (rc) return rc; R.ext_lo[k] = R.lo[k]; R.ext_hi[k] = R.hi[k]; }
      if (R.lo[0] > R.hi[0] || R.lo[1] > R.hi[1] || R.lo[2] > R.hi[2]) return fail(d, ERR_ARG, "Illegal region block command");
    }
    for (size_t k = opt; k < w.size(); k += 2) {  // Region::options (region.cpp:380-460)
      if (k + 1 >= w.size()) return fail(d, ERR_ARG, "Illegal region command");
      if (w[k] == "units") { if (w[k + 1] != "box") return fail(d, ERR_UNSUPPORTED, "region units lattice is outside the hot-path scope"); }
      else if (w[k] == "seed") { int sd; rc = take_seed(d, w[k + 1], sd); if (rc) return rc; R.random.reset(sd); }
      else if (w[k] == "side" && w[k + 1] == "in") {}
      else return fail(d, ERR_UNSUPPORTED, "region keyword '%s' is outside the hot-path scope", w[k].c_str());
    }
    d->regions[w[1]] = R; return OK;
  }
  if (c == "create_box") {
    if (w.size() != 3) return fail(d, ERR_ARG, "Illegal create_box command");
    if (d->have_box) return fail(d, ERR_ARG, "Cannot create_box after simulation box is defined");
    if (d->opaque_regions.count(w[2])) return fail(d, ERR_UNSUPPORTED, "create_box from a region that is not a block is outside the hot-path scope");
    if (!d->regions.count(w[2])) return fail(d, ERR_ARG, "Create_box region ID does not exist");
    if (d->regions[w[2]].kind != 0) return fail(d, ERR_UNSUPPORTED, "create_box from a region that is not a block is outside the hot-path scope");
    rc = inumeric(d, w[1], d->ntypes); if (rc) return rc;
    for (int k = 0; k < 3; k++) { d->lo[k] = d->regions[w[2]].lo[k]; d->hi[k] = d->regions[w[2]].hi[k]; }
    d->have_box = true; return send_box(d);
  }
  if (c == "read_data") { if (w.size() < 2) return fail(d, ERR_ARG, "Illegal read_data command"); rc = read_data(d, w[1]); if (rc) return rc; return send_box(d); }
  if (c == "neighbor") {
    if (w.size() != 3) return fail(d, ERR_ARG, "Illegal neighbor command");
    if (w[2] != "bin") return fail(d, ERR_UNSUPPORTED, "neighbor style %s is outside the hot-path scope (bin)", w[2].c_str());
    rc = numeric(d, w[1], d->skin); if (rc) return rc;
    TRY(API(set_neighbor)(d->e, d->skin, d->every, d->delay, d->check)); return OK;
  }
  if (c == "neigh_modify") {  // neighbor.cpp:2180-2290
    for (size_t k = 1; k < w.size(); k += 2) {
      if (k + 1 >= w.size()) return fail(d, ERR_ARG, "Illegal neigh_modify command");
      if (w[k] == "every") { rc = inumeric(d, w[k + 1], d->every); if (rc) return rc; }
      else if (w[k] == "delay") { rc = inumeric(d, w[k + 1], d->delay); if (rc) return rc; }
      else if (w[k] == "check") { if (w[k + 1] != "yes" && w[k + 1] != "no") return fail(d, ERR_ARG, "Illegal neigh_modify command"); d->check = w[k + 1] == "yes"; }
      else if (w[k] == "contact_distance_factor") { double f; rc = numeric(d, w[k + 1], f); if (rc) return rc; TRY(API(set_contact_distance_factor)(d->e, f)); }
      else if (w[k] == "page" || w[k] == "one" || w[k] == "binsize") continue;
      else return fail(d, ERR_UNSUPPORTED, "neigh_modify keyword '%s' is outside the hot-path scope", w[k].c_str());
    }
    TRY(API(set_neighbor)(d->e, d->skin, d->every, d->delay, d->check)); return OK;
  }
  if (c == "pair_style") {
    if (w.size() < 2 || w[1] != "gran") return fail(d, ERR_UNSUPPORTED, "pair_style %s is outside the hot-path scope (gran)", w.size() > 1 ? w[1].c_str() : "");
    rc = send_box(d); if (rc) return rc;
    std::vector<const char *> a = cptrs(w, 2);
    TRY(API(set_pair_style)(d->e, (int)a.size(), a.data()));
    d->have_pair = true; return OK;
  }
  if (c == "fix") return cmd_fix(d, w);
  if (c == "include") {  // Input::include (input.cpp:640-670): the named script runs in place; relative to the working directory like
    // every file name of a deck -- here: to the directory of the deck that is being read
    if (w.size() != 2) return fail(d, ERR_ARG, "Illegal include command");
    std::string path = w[1];
    if (!path.empty() && path[0] != '/' && !d->dir.empty()) path = d->dir + "/" + path;
    const std::string keep = d->dir;
    rc = run_file(d, path);
    d->dir = keep;
    return rc;
  }
  if (c == "unfix") {
    if (w.size() != 2) return fail(d, ERR_ARG, "Illegal unfix command");
    if (d->ignored_fixes.erase(w[1])) return OK;
    for (size_t k = 0; k < d->inserts.size(); k++) if (d->inserts[k].id == w[1]) { d->inserts.erase(d->inserts.begin() + k); return OK; }
    if (d->extra_fixes.count(w[1])) { TRY(API(set_extra_force)(d->e, w[1].c_str(), 0, 0, nullptr, -1)); d->extra_fixes.erase(w[1]); return OK; }
    return fail(d, ERR_UNSUPPORTED, "unfix of a hot-path fix is outside the hot-path scope");
  }
  if (c == "group") return cmd_group(d, w);
  if (c == "timestep") { if (w.size() != 2) return fail(d, ERR_ARG, "Illegal timestep command"); double dt; rc = numeric(d, w[1], dt); if (rc) return rc; TRY(API(set_timestep)(d->e, dt)); d->dt = dt; return OK; }
  if (c == "thermo") {  // thermo N (output.cpp:430-460)
    if (w.size() != 2) return fail(d, ERR_ARG, "Illegal thermo command");
    double nd; rc = numeric(d, w[1], nd); if (rc) return rc;
    if ((long)nd < 0) return fail(d, ERR_ARG, "Illegal thermo command");
    d->thermo_every = (long)nd; return OK;
  }
  if (c == "thermo_style") {  // thermo_style one | custom kw ... : columns outside the path's state (computes, variables, fixes) are dropped with a note
    if (w.size() < 2) return fail(d, ERR_ARG, "Illegal thermo_style command");
    if (w[1] == "one" || w[1] == "multi") { d->thermo_kw = {"step", "atoms", "ke", "cpu"}; return OK; }
    if (w[1] != "custom") return fail(d, ERR_ARG, "Illegal thermo style command");
    d->thermo_kw.clear();
    for (size_t k = 2; k < w.size(); k++) {
      if (thermo_col(w[k])) d->thermo_kw.push_back(w[k]);
      else d->warnings += "thermo keyword " + w[k] + " ignored (output only)\n";
    }
    return OK;
  }
  if (c == "dump") {  // dump ID group custom N file field ...   (dump.cpp:60-110, dump_custom.cpp:95-180)
    if (w.size() < 6) return fail(d, ERR_ARG, "Illegal dump command");
    if (w[3] != "custom") { d->warnings += "dump style " + w[3] + " ignored (output only)\n"; return OK; }
    if (w[2] != "all") return fail(d, ERR_UNSUPPORTED, "dump custom: only group 'all' is on the hot path");
    if (w.size() == 6) return fail(d, ERR_ARG, "No dump custom arguments specified");
    double nd; rc = numeric(d, w[4], nd); if (rc) return rc;
    if ((long)nd <= 0) return fail(d, ERR_ARG, "Invalid dump frequency");
    Deck::Dump D; D.every = (long)nd; D.file = w[5]; D.fields.assign(w.begin() + 6, w.end());
    d->dumps[w[1]] = D; return OK;
  }
  if (c == "undump") { if (w.size() != 2) return fail(d, ERR_ARG, "Illegal undump command"); d->dumps.erase(w[1]); return OK; }
  if (c == "dump_modify") {  // atoms are always written in ascending id (== `sort id`); other keywords do not change the content
    if (w.size() < 2) return fail(d, ERR_ARG, "Illegal dump_modify command");
    for (size_t k = 2; k < w.size(); k++) if (w[k] == "format") return fail(d, ERR_UNSUPPORTED, "dump_modify format is outside the hot-path scope");
    if (d->dumps.count(w[1])) for (size_t k = 2; k + 1 < w.size(); k++) {
      if (w[k] == "pad") { rc = inumeric(d, w[k + 1], d->dumps[w[1]].pad); if (rc) return rc; }
      else if (w[k] == "first") d->dumps[w[1]].first = (w[k + 1] == "yes");
    }
    return OK;
  }
  if (c == "lattice") {  // lattice.cpp:60-300
    if (w.size() < 2) return fail(d, ERR_ARG, "Illegal lattice command");
    Deck::Lattice L;
    if (w[1] == "none") { d->lattice = L; return OK; }
    if (w.size() < 3) return fail(d, ERR_ARG, "Illegal lattice command");
    if (w[1] == "sc") L.basis = {0., 0., 0.};
    else if (w[1] == "bcc") L.basis = {0., 0., 0., .5, .5, .5};
    else if (w[1] == "fcc") L.basis = {0., 0., 0., .5, .5, 0., .5, 0., .5, 0., .5, .5};
    else return fail(d, ERR_UNSUPPORTED, "lattice style '%s' is outside the hot-path scope (sc, bcc, fcc)", w[1].c_str());
    rc = numeric(d, w[2], L.scale); if (rc) return rc;
    if (L.scale <= 0.0) return fail(d, ERR_ARG, "Illegal lattice command");
    for (size_t k = 3; k < w.size();) {
      if (w[k] == "origin" && k + 3 < w.size()) {
        for (int q = 0; q < 3; q++) { rc = numeric(d, w[k + 1 + q], L.origin[q]); if (rc) return rc; if (L.origin[q] < 0.0 || L.origin[q] >= 1.0) return fail(d, ERR_ARG, "Illegal lattice command"); }
        k += 4;
      } else return fail(d, ERR_UNSUPPORTED, "lattice keyword '%s' is outside the hot-path scope", w[k].c_str());
    }
    L.defined = true; d->lattice = L; return OK;
  }
  if (c == "create_atoms" && w.size() >= 3 && (w[2] == "box" || w[2] == "region")) {  // create_atoms.cpp:100-300, add_lattice :492-602
    if (!d->have_box) return fail(d, ERR_STATE, "Create_atoms command before simulation box is defined");
    if (!d->lattice.defined) return fail(d, ERR_ARG, "Cannot create atoms with undefined lattice");
    double t; rc = numeric(d, w[1], t); if (rc) return rc;
    if ((int)t < 1 || (int)t > d->ntypes) return fail(d, ERR_ARG, "Invalid atom type in create_atoms command");
    const Deck::Region *R = nullptr;
    size_t opt = 3;
    if (w[2] == "region") {
      if (w.size() < 4) return fail(d, ERR_ARG, "Illegal create_atoms command");
      if (d->opaque_regions.count(w[3])) return fail(d, ERR_UNSUPPORTED, "create_atoms in a region that is neither a block nor a cylinder is outside the hot-path scope");
      if (!d->regions.count(w[3])) return fail(d, ERR_ARG, "Create_atoms region ID does not exist");
      R = &d->regions[w[3]]; opt = 4;
    }
    for (size_t k = opt; k < w.size(); k += 2) {
      if (w[k] == "units" && k + 1 < w.size()) continue;  // (only the style single reads it)
      return fail(d, ERR_UNSUPPORTED, "create_atoms keyword '%s' is outside the hot-path scope", w[k].c_str());
    }
    // my sub-box = the whole box; a periodic dimension is shifted down by EPSILON of its length so that exactly one of two
    // images on the boundary is created (create_atoms.cpp:205-262)
    double sublo[3], subhi[3];
    for (int k = 0; k < 3; k++) {
      sublo[k] = d->lo[k]; subhi[k] = d->hi[k];
      if (d->periodic[k]) { const double eps = (d->hi[k] - d->lo[k]) * 1.0e-6; sublo[k] -= eps; subhi[k] -= 2.0 * eps; }
    }
    const Deck::Lattice &L = d->lattice;
    int lo[3], hi[3];  // cells that can reach the box (the reference's bounds are tighter; cells outside create nothing)
    for (int k = 0; k < 3; k++) { lo[k] = (int)std::floor(d->lo[k] / L.scale) - 2; hi[k] = (int)std::ceil(d->hi[k] / L.scale) + 2; }
    const bool pend = d->uploaded;
    for (int v : d->tag) d->maxtag = std::max(d->maxtag, v);
    const int nb = (int)L.basis.size() / 3;
    for (int k = lo[2]; k <= hi[2]; k++) for (int j = lo[1]; j <= hi[1]; j++) for (int i = lo[0]; i <= hi[0]; i++) for (int m = 0; m < nb; m++) {
      double x[3] = {i + L.basis[3 * m], j + L.basis[3 * m + 1], k + L.basis[3 * m + 2]};
      for (int q = 0; q < 3; q++) { x[q] *= L.scale; x[q] = x[q] + L.scale * L.origin[q]; }  // Lattice::lattice2box with unit primitive / rotation matrices
      if (R && !region_inside(*R, x[0], x[1], x[2])) continue;
      if (x[0] < sublo[0] || x[0] >= subhi[0] || x[1] < sublo[1] || x[1] >= subhi[1] || x[2] < sublo[2] || x[2] >= subhi[2]) continue;
      const int id = ++d->maxtag;
      (pend ? d->ptag : d->tag).push_back(id); (pend ? d->ptype : d->type).push_back((int)t); (pend ? d->pmask : d->mask).push_back(1);
      for (int q = 0; q < 3; q++) { (pend ? d->px : d->x).push_back(x[q]); (pend ? d->pv : d->v).push_back(0.0); (pend ? d->pomega : d->omega).push_back(0.0); }
      (pend ? d->pradius : d->radius).push_back(0.5); (pend ? d->pdensity : d->density).push_back(1.0);
    }
    return OK;
  }
  if (c == "create_atoms") {  // create_atoms.cpp (style single): one atom of the given type at a point; radius 0.5, density 1 until `set` (atom_vec_sphere.cpp:167-190)
    if (w.size() < 6 || w[2] != "single") return fail(d, ERR_UNSUPPORTED, "create_atoms: styles 'single', 'box' and 'region' are on the hot path (random creation is outside its scope)");
    if (!d->have_box) return fail(d, ERR_STATE, "Create_atoms command before simulation box is defined");
    double t, p[3];
    rc = numeric(d, w[1], t); if (rc) return rc;
    for (int k = 0; k < 3; k++) { rc = numeric(d, w[3 + k], p[k]); if (rc) return rc; }
    for (size_t k = 6; k < w.size(); k++) {
      if (w[k] == "units" && k + 1 < w.size() && w[k + 1] == "box") k++;
      else return fail(d, ERR_UNSUPPORTED, "create_atoms keyword '%s' is outside the hot-path scope", w[k].c_str());
    }
    if ((int)t < 1 || (int)t > d->ntypes) return fail(d, ERR_ARG, "Invalid atom type in create_atoms command");
    for (int v : d->tag) d->maxtag = std::max(d->maxtag, v);
    const int id = ++d->maxtag;
    const bool pend = d->uploaded;
    (pend ? d->ptag : d->tag).push_back(id); (pend ? d->ptype : d->type).push_back((int)t); (pend ? d->pmask : d->mask).push_back(1);
    for (int k = 0; k < 3; k++) { (pend ? d->px : d->x).push_back(p[k]); (pend ? d->pv : d->v).push_back(0.0); (pend ? d->pomega : d->omega).push_back(0.0); }
    (pend ? d->pradius : d->radius).push_back(0.5); (pend ? d->pdensity : d->density).push_back(1.0);
    return OK;
  }
  if (c == "velocity") {  // velocity.cpp:160-260 (style set; LIGGGHTS default `units box`): initial velocity of the atoms of a group
    if (w.size() < 6 || w[2] != "set") return fail(d, ERR_UNSUPPORTED, "velocity: only style 'set' is on the hot path");
    if (!d->groups.count(w[1])) return fail(d, ERR_ARG, "Could not find velocity group ID");
    const int bit = d->groups[w[1]];
    bool sum = false;
    for (size_t k = 6; k + 1 < w.size(); k += 2) {
      if (w[k] == "sum") sum = w[k + 1] == "yes";
      else if (w[k] == "units") { if (w[k + 1] != "box") return fail(d, ERR_UNSUPPORTED, "velocity units lattice is outside the hot-path scope"); }
      else return fail(d, ERR_UNSUPPORTED, "velocity keyword '%s' is outside the hot-path scope", w[k].c_str());
    }
    const bool pend = d->uploaded;
    std::vector<int> &mk = pend ? d->pmask : d->mask;
    std::vector<double> &vv = pend ? d->pv : d->v;
    if (pend && mk.empty()) return fail(d, ERR_UNSUPPORTED, "velocity set: the atoms are already on the engine (velocities of running particles cannot be changed)");
    for (int k = 0; k < 3; k++) {
      if (w[3 + k] == "NULL") continue;
      double val; rc = numeric(d, w[3 + k], val); if (rc) return rc;
      for (size_t i = 0; i < mk.size(); i++) if (mk[i] & bit) vv[3 * i + k] = sum ? vv[3 * i + k] + val : val;
    }
    return OK;
  }
  if (c == "set") {  // set.cpp, styles atom and group: diameter / density / type of atoms that have not reached the engine yet
    if (w.size() < 5 || (w[1] != "atom" && w[1] != "group")) return fail(d, ERR_UNSUPPORTED, "set: only styles 'atom' and 'group' are on the hot path");
    int lo = 1, hi = 0x7fffffff, gbit = 0;
    if (w[1] == "group") { if (!d->groups.count(w[2])) return fail(d, ERR_ARG, "Could not find set group ID"); gbit = d->groups[w[2]]; }
    else { const std::string &r = w[2]; const size_t star = r.find('*');
      if (star == std::string::npos) { double a; rc = numeric(d, r, a); if (rc) return rc; lo = hi = (int)a; }
      else { lo = star ? atoi(r.substr(0, star).c_str()) : 1; hi = star + 1 < r.size() ? atoi(r.substr(star + 1).c_str()) : 0x7fffffff; } }
    const bool pend = d->uploaded;
    std::vector<int> &tg = pend ? d->ptag : d->tag, &ty = pend ? d->ptype : d->type, &mk = pend ? d->pmask : d->mask;
    std::vector<double> &ra = pend ? d->pradius : d->radius, &de = pend ? d->pdensity : d->density;
    bool any = false;
    for (size_t i = 0; i < tg.size(); i++) {
      if (gbit ? !(mk[i] & gbit) : (tg[i] < lo || tg[i] > hi)) continue;
      any = true;
      for (size_t k = 3; k + 1 < w.size(); k += 2) {
        double val; rc = numeric(d, w[k + 1], val); if (rc) return rc;
        if (w[k] == "diameter") { if (!(val > 0)) return fail(d, ERR_ARG, "Invalid diameter in set command"); ra[i] = 0.5 * val; }
        else if (w[k] == "density") { if (!(val > 0)) return fail(d, ERR_ARG, "Invalid density in set command"); de[i] = val; }
        else if (w[k] == "type") { if ((int)val < 1 || (int)val > d->ntypes) return fail(d, ERR_ARG, "Invalid value in set command"); ty[i] = (int)val; }
        else return fail(d, ERR_UNSUPPORTED, "set keyword '%s' is outside the hot-path scope", w[k].c_str());
      }
    }
    if (!any) return fail(d, ERR_UNSUPPORTED, "set atom: the atoms are already on the engine (properties of running particles cannot be changed)");
    return OK;
  }
  if (c == "run") {  // run.cpp:40-130: every run starts with Verlet::setup
    if (w.size() < 2) return fail(d, ERR_ARG, "Illegal run command");
    double nd; rc = numeric(d, w[1], nd); if (rc) return rc;
    long n = (long)nd;
    for (size_t k = 2; k < w.size(); k++) {
      if (w[k] == "upto") { n -= d->ntimestep; if (n < 0) return fail(d, ERR_ARG, "Run command upto value is before current timestep"); }
      else return fail(d, ERR_UNSUPPORTED, "run keyword '%s' is outside the hot-path scope", w[k].c_str());
    }
    if (n < 0) return fail(d, ERR_ARG, "Invalid run command N value");
    rc = first_run_prepare(d); if (rc) return rc;
    if (d->uploaded) TRY(API(setup)(d->e));  // (else the box is still empty: nothing to set up)
    for (auto &I : d->inserts) { rc = insert_setup(d, I); if (rc) return rc; }  // Modify::setup -> FixInsert::setup
    rc = write_dumps_due(d, true); if (rc) return rc;
    const long run_first = d->ntimestep;
    d->loop0 = std::chrono::steady_clock::now();
    thermo_header(d);
    rc = thermo_line(d, true, run_first); if (rc) return rc;
    while (n > 0) {  // the run in slices that end on the next output step (no setup in between: one run of the reference)
      long chunk = n;
      for (auto &kv : d->dumps) if (kv.second.every > 0) chunk = std::min(chunk, (d->ntimestep / kv.second.every + 1) * kv.second.every - d->ntimestep);
      if (d->thermo_every > 0) chunk = std::min(chunk, (d->ntimestep / d->thermo_every + 1) * d->thermo_every - d->ntimestep);
      long next_ins = -1;  // the next timestep a fix insert/pack acts on (fix->next_reneighbor)
      for (auto &I : d->inserts) if (I.next > d->ntimestep && (next_ins < 0 || I.next < next_ins)) next_ins = I.next;
      if (next_ins == d->ntimestep + 1) { rc = insertion_step(d); if (rc) return rc; chunk = 1; }
      else {
        if (next_ins > 0) chunk = std::min(chunk, next_ins - 1 - d->ntimestep);
        if (d->uploaded) TRY(API(run)(d->e, chunk));  // (an empty box has nothing to step)
      }
      d->ntimestep += chunk; n -= chunk;
      rc = write_dumps_due(d); if (rc) return rc;
      if (n == 0 || (d->thermo_every > 0 && d->ntimestep % d->thermo_every == 0)) { rc = thermo_line(d, false, run_first); if (rc) return rc; }  // (thermo.cpp: every N steps and on the last step)
    }
    return OK;
  }
  return fail(d, ERR_UNSUPPORTED, "command '%s' is outside the hot-path scope", c.c_str());
}

// a whole input script, '&' continuation lines joined (Input::file input.cpp:160-250); stops at the first error.  File names
// inside the script stay relative to the directory of the outermost script (the reference's working directory).
int run_file(Deck *d, const std::string &path)
{
  std::ifstream f(path);
  if (!f.good()) return fail(d, ERR_ARG, "Cannot open input script %s", path.c_str());
  std::string line, acc; int lineno = 0, rc = OK;
  while (std::getline(f, line)) {
    lineno++;
    size_t end = line.find_last_not_of(" \t\r\n");
    if (end != std::string::npos && line[end] == '&') { acc += line.substr(0, end) + " "; continue; }
    acc += line;
    try { rc = one(d, acc); } catch (const std::exception &ex) { d->err = ex.what(); rc = ERR_ARG; }
    if (rc != OK) { char where[64]; snprintf(where, sizeof where, " (line %d)", lineno); d->err += where; break; }
    acc.clear();
  }
  return rc;
}

}  // namespace

extern "C" {

typedef struct DECK(handle) { Deck d; } DECK(handle);

// `lammps_open_no_mpi` analogue: the deck state wraps an engine the caller created (and still owns)
int DECK(open)(DECK(handle) **out, dem_engine *e)
{
  if (!out || !e) return ERR_ARG;
  DECK(handle) *h = new DECK(handle)();
  h->d.e = e; h->d.groups["all"] = 1;
  *out = h; return OK;
}
void DECK(close)(DECK(handle) *h) { delete h; }
const char *DECK(last_error)(const DECK(handle) *h) { return h ? h->d.err.c_str() : "null deck"; }
const char *DECK(warnings)(const DECK(handle) *h) { return h ? h->d.warnings.c_str() : ""; }
long DECK(ntimestep)(const DECK(handle) *h) { return h ? h->d.ntimestep : -1; }
// thermo output (header + lines) collected so far; screen(1): also print it to stdout as it is produced (`lmp_b200`)
const char *DECK(output)(const DECK(handle) *h) { return h ? h->d.out.c_str() : ""; }
int DECK(screen)(DECK(handle) *h, int on) { if (!h) return ERR_ARG; h->d.screen = on != 0; return OK; }
// `lammps_command`: one input-script line (no continuation handling, as Input::one)
int DECK(command)(DECK(handle) *h, const char *line)
{
  if (!h || !line) return ERR_ARG;
  try { return one(&h->d, line); } catch (const std::exception &ex) { h->d.err = ex.what(); return ERR_ARG; }
}
// `lammps_file`: a whole input script, '&' continuation lines joined (Input::file input.cpp:160-250); stops at the first error
int DECK(file)(DECK(handle) *h, const char *path)
{
  if (!h || !path) return ERR_ARG;
  const std::string p(path); const size_t sl = p.rfind('/');
  const std::string olddir = h->d.dir;
  h->d.dir = sl == std::string::npos ? "" : p.substr(0, sl);
  const int rc = run_file(&h->d, p);
  h->d.dir = olddir;
  return rc;
}

}  // extern "C"
