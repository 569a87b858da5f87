// On-disk formats either side of the path: host code, no CUDA.
//
//  * the failed-node dump of TMINLP2TNLP::outputDiffs (src/Interfaces/BonTMINLP2TNLP.cpp:551-615): `<name>.bounds`
//    (one "LO|UP <tab> index <tab> value" line per bound that differs from the original model) and `<name>.start`
//    ("numcols <tab> 2 numcols + numrows", then the primal start point and, when present, the duals, 17 digits);
//  * their readers BoundsReader::read/readAndApply (src/Interfaces/BonBoundsReader.cpp:52-119) and
//    StartPointReader::readFile (src/Interfaces/BonStartPointReader.cpp:15-37) — what Apps/BonNodeSolver.cpp:72-81 replays;
//  * a BATCH replay file: the children of one bnlp_solve_batch / bnlp_solve_children call in the same line format, one
//    "CHILD k" section per child, so that a GPU/CPU disagreement on child k becomes a two-file repro (`.children` +
//    `.start`) that either the engine, the oracle or (child by child, through the .bounds form) the reference can replay;
//  * the AMPL solution file written by AmplTMINLP::finalize_solution / write_solution
//    (src/Interfaces/Ampl/BonAmplTMINLP.cpp:712-794).  The writer the reference calls, ASL's write_sol, is a third-party
//    dependency that is not in the reference tree (ThirdParty/ASL, fetched by get.ASL): its published text format is
//    restated here — message, blank line, "Options" block echoing the option integers of the .nl header, the four counts
//    (constraints, duals written, variables, primals written), the values, "objno 0 <solve_result_num>" — PARITY UNPINNED
//    against ASL itself; the non-AMPL fallback of :758-766 ("i <tab> x_i" lines, then -1) is the reference's own code.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include <charconv>

#include "../../include/bnlp.h"

namespace {

// CoinRelFltEq with its default epsilon 1e-10 (what outputDiffs compares bounds with)
bool rel_eq(double a, double b) {
  if (a == b) return true;
  if (std::isnan(a) || std::isnan(b)) return false;
  const double tol = (std::fabs(a) > std::fabs(b) ? std::fabs(a) : std::fabs(b)) + 1.0;
  return std::fabs(a - b) <= 1e-10 * tol;
}

// shortest decimal string that reads back to the same double (what ASL's g_fmt produces)
std::string shortest(double v) {
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof(buf), v);
  return std::string(buf, r.ptr);
}

void write_bound_lines(std::ostream &f, int n, const double *xl, const double *xu, const double *oxl, const double *oxu) {
  for (int i = 0; i < n; ++i) {
    if (!rel_eq(xl[i], oxl[i])) f << "LO" << "\t" << i << "\t" << xl[i] << std::endl;
    if (!rel_eq(xu[i], oxu[i])) f << "UP" << "\t" << i << "\t" << xu[i] << std::endl;
  }
}

}  // namespace

extern "C" {

int bnlp_write_node_files(const char *prob_name, int n, int m, const double *xl, const double *xu, const double *orig_xl,
                          const double *orig_xu, const double *x_init, const double *duals_init) {
  if (!prob_name || !xl || !xu || !orig_xl || !orig_xu || !x_init) return -1;
  std::ofstream fb(std::string(prob_name) + ".bounds");
  if (!fb) return -1;
  write_bound_lines(fb, n, xl, xu, orig_xl, orig_xu);
  std::ofstream fs(std::string(prob_name) + ".start");
  if (!fs) return -1;
  fs.precision(17);
  fs << n << "\t" << 2 * n + m << std::endl;
  for (int i = 0; i < n; ++i) fs << x_init[i] << std::endl;
  if (duals_init)
    for (int i = 0; i < 2 * n + m; ++i) fs << duals_init[i] << std::endl;
  return (fb && fs) ? 0 : -1;
}

int bnlp_read_bounds_file(const char *file, int n, double *xl, double *xu) {
  std::ifstream f(file);
  if (!f) return -1;
  std::string tag;
  int applied = 0;
  while (f >> tag) {
    int idx;
    double b;
    if (!(f >> idx >> b)) return -1;
    if (idx < 0 || idx >= n) return -1;
    if (tag == "LO") xl[idx] = b;
    else if (tag == "UP") xu[idx] = b;
    else return -1;   // the reference throws on any other tag
    f.ignore(10000, '\n');
    ++applied;
  }
  return applied;
}

int bnlp_read_start_file(const char *file, int *n_primals, int *n_duals, double *primals, int cap_primals, double *duals,
                         int cap_duals) {
  std::ifstream f(file);
  if (!f) return -1;
  int np = 0, nd = 0;
  if (!(f >> np >> nd) || np < 0 || nd < 0) return -1;
  if (n_primals) *n_primals = np;
  if (n_duals) *n_duals = nd;
  int got_d = 0;
  for (int i = 0; i < np; ++i) {
    double v;
    if (!(f >> v)) return -1;
    if (primals && i < cap_primals) primals[i] = v;
  }
  for (int i = 0; i < nd; ++i) {   // the dump omits the duals when there are none: report how many were present
    double v;
    if (!(f >> v)) break;
    if (duals && i < cap_duals) duals[i] = v;
    ++got_d;
  }
  if (n_duals) *n_duals = got_d;
  return 0;
}

int bnlp_write_batch_replay(const char *prob_name, int n, int m, int B, const double *XL, const double *XU,
                            const double *orig_xl, const double *orig_xu, const double *X0, int x0_per_child) {
  if (!prob_name || !XL || !XU || !orig_xl || !orig_xu || !X0 || B < 0) return -1;
  std::ofstream fc(std::string(prob_name) + ".children");
  if (!fc) return -1;
  fc.precision(17);   // (the single-node dump keeps the stream default of 6 digits, as the reference; a replay must be exact)
  fc << "BATCH" << "\t" << n << "\t" << m << "\t" << B << "\t" << (x0_per_child ? 1 : 0) << std::endl;
  for (int b = 0; b < B; ++b) {
    fc << "CHILD" << "\t" << b << std::endl;
    write_bound_lines(fc, n, XL + (size_t)b * n, XU + (size_t)b * n, orig_xl, orig_xu);
  }
  std::ofstream fs(std::string(prob_name) + ".start");
  if (!fs) return -1;
  fs.precision(17);
  const int rows = x0_per_child ? B : 1;
  fs << rows * n << "\t" << 0 << std::endl;
  for (size_t i = 0; i < (size_t)rows * n; ++i) fs << X0[i] << std::endl;
  return (fc && fs) ? 0 : -1;
}

int bnlp_read_batch_replay_header(const char *file, int *n, int *m, int *B, int *x0_per_child) {
  std::ifstream f(file);
  std::string tag;
  int a, b, c, d;
  if (!f || !(f >> tag >> a >> b >> c >> d) || tag != "BATCH") return -1;
  if (n) *n = a;
  if (m) *m = b;
  if (B) *B = c;
  if (x0_per_child) *x0_per_child = d;
  return 0;
}

int bnlp_read_batch_replay(const char *file, int n, int B, const double *orig_xl, const double *orig_xu, double *XL,
                           double *XU) {
  std::ifstream f(file);
  std::string tag;
  int hn, hm, hB, hx;
  if (!f || !(f >> tag >> hn >> hm >> hB >> hx) || tag != "BATCH" || hn != n || hB != B) return -1;
  for (int b = 0; b < B; ++b) {
    std::memcpy(XL + (size_t)b * n, orig_xl, sizeof(double) * n);
    std::memcpy(XU + (size_t)b * n, orig_xu, sizeof(double) * n);
  }
  int cur = -1;
  while (f >> tag) {
    if (tag == "CHILD") {
      if (!(f >> cur) || cur < 0 || cur >= B) return -1;
      continue;
    }
    int idx;
    double v;
    if (cur < 0 || !(f >> idx >> v) || idx < 0 || idx >= n) return -1;
    if (tag == "LO") XL[(size_t)cur * n + idx] = v;
    else if (tag == "UP") XU[(size_t)cur * n + idx] = v;
    else return -1;
  }
  return 0;
}

// AmplTMINLP::finalize_solution (BonAmplTMINLP.cpp:712-753): message and AMPL solve_result_num per TMINLP::SolverReturn
// (0 SUCCESS, 1 INFEASIBLE, 2 CONTINUOUS_UNBOUNDED, 3 LIMIT_EXCEEDED, 4 USER_INTERRUPT, 5 MINLP_ERROR)
int bnlp_sol_status(int solver_return, int has_x, const char *app_name, char *message, size_t msglen, int *solve_result_num,
                    int *finished) {
  const std::string app = app_name ? app_name : "bonmin";
  std::string msg;
  int num = 500, fin = 0;
  switch (solver_return) {
    case 0: msg = "\n" + app + ": Optimal"; num = 3; fin = 1; break;
    case 1: msg = "\n" + app + ": Infeasible problem"; num = 220; fin = 1; break;
    case 2: msg = "\n" + app + " Continuous relaxation is unbounded."; num = 300; fin = 1; break;
    case 3: msg = "\n" + app + ": Optimization interrupted on limit."; num = has_x ? 421 : 410; break;
    case 4: msg = "\n" + app + ": Optimization interrupted by user."; num = has_x ? 422 : 411; break;
    case 5: msg = "\n" + app + ": Error encountered in optimization."; num = 500; fin = -1; break;
    default: return -1;
  }
  if (message && msglen) { std::strncpy(message, msg.c_str(), msglen - 1); message[msglen - 1] = 0; }
  if (solve_result_num) *solve_result_num = num;
  if (finished) *finished = fin;   // 1 "Finished", 0 "Not finished", -1 "Aborted"
  return 0;
}

// option integers of the first line of a text .nl file ("g3 1 1 0" -> 3 1 1 0): what write_sol echoes under "Options"
int bnlp_nl_header_options(const char *nl_text, size_t len, int *opts, int cap, double *vbtol) {
  if (!nl_text || len < 2 || (nl_text[0] != 'g' && nl_text[0] != 'b')) return -1;
  size_t e = 0;
  while (e < len && nl_text[e] != '\n') ++e;
  std::istringstream is(std::string(nl_text + 1, e - 1));
  int cnt = 0, v;
  if (!(is >> cnt) || cnt < 0 || cnt > 9) return -1;
  if (opts && cap > 0) opts[0] = cnt;
  int got = 0;
  while (got < cnt && (is >> v)) {
    if (opts && got + 1 < cap) opts[got + 1] = v;
    ++got;
  }
  if (got != cnt) return -1;
  double vb = 0.0;
  if (cnt >= 2 && opts && cap > 2 && opts[2] == 3) is >> vb;   // ampl_vbtol follows the integers when option 2 is 3
  if (vbtol) *vbtol = vb;
  return cnt;
}

int bnlp_write_sol(const char *file, const char *message, const int *ampl_options, double ampl_vbtol, int n_con, int n_var,
                   const double *x, const double *y, int solve_result_num) {
  if (!file || !message) return -1;
  std::ofstream f(file);
  if (!f) return -1;
  f << message << "\n\n" << "Options\n";
  const int cnt = ampl_options ? ampl_options[0] : 0;
  f << cnt << "\n";
  for (int i = 1; i <= cnt; ++i) f << ampl_options[i] << "\n";
  if (cnt >= 2 && ampl_options[2] == 3) f << shortest(ampl_vbtol) << "\n";
  f << n_con << "\n" << (y ? n_con : 0) << "\n" << n_var << "\n" << (x ? n_var : 0) << "\n";
  if (y) for (int i = 0; i < n_con; ++i) f << shortest(y[i]) << "\n";
  if (x) for (int i = 0; i < n_var; ++i) f << shortest(x[i]) << "\n";
  f << "objno 0 " << solve_result_num << "\n";
  return f ? 0 : -1;
}

// the non-AMPL fallback of finalize_solution (:758-766): "<appName>.sol" with "i <tab> x_i" lines and a closing -1
int bnlp_write_plain_sol(const char *file, int n, const double *x) {
  std::ofstream f(file);
  if (!f) return -1;
  for (int i = 0; i < n; ++i) f << i << "\t" << x[i] << std::endl;
  f << "-1\n";
  return f ? 0 : -1;
}

}  // extern "C"
