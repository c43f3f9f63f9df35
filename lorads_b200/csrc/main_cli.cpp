// main_cli.cpp -- stand-alone command-line driver on top of the C ABI: `lorads_b200_cli file.dat-s [--options]`.
// The option names and defaults are the reference driver's (src_semi/main.c:19-43,57-80,168-236); the phase
// sequence itself (main.c:321-487) lives in lb2_solve.  The reference's own unmodified main.c can be linked
// against the library instead (integration/lorads_dropin.c); this driver exists so that the GPU path is usable
// without the reference tree.
#include <getopt.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/lorads_b200.h"

static struct option long_options[] = {
    {"initRho", required_argument, 0, 1000},      {"rhoMax", required_argument, 0, 1001},
    {"rhoCellingALM", required_argument, 0, 1002}, {"rhoCellingADMM", required_argument, 0, 1003},
    {"maxALMIter", required_argument, 0, 1004},   {"maxADMMIter", required_argument, 0, 1005},
    {"timesLogRank", required_argument, 0, 1006}, {"rhoFreq", required_argument, 0, 1007},
    {"rhoFactor", required_argument, 0, 1008},    {"ALMRhoFactor", required_argument, 0, 1009},
    {"phase1Tol", required_argument, 0, 1010},    {"phase2Tol", required_argument, 0, 1011},
    {"timeSecLimit", required_argument, 0, 1012}, {"heuristicFactor", required_argument, 0, 1013},
    {"lbfgsListLength", required_argument, 0, 1014}, {"endTauTol", required_argument, 0, 1015},
    {"endALMSubTol", required_argument, 0, 1016}, {"l2Rescaling", required_argument, 0, 1017},
    {"reoptLevel", required_argument, 0, 1018},   {"dyrankLevel", required_argument, 0, 1019},
    {"highAccMode", required_argument, 0, 1020},  {"device", required_argument, 0, 2000},
    {"quiet", no_argument, 0, 2001},              {"saveBinary", required_argument, 0, 2002},
    {0, 0, 0, 0}};

static void fail(const char *what) {
    std::fprintf(stderr, "lorads_b200_cli: %s: %s\n", what, lb2_last_error());
    std::exit(2);
}

int main(int argc, char **argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s file.dat-s [--timesLogRank x] [--phase2Tol x] [--timeSecLimit x] ... [--device k] [--quiet] [--saveBinary file]\n", argv[0]);
        return 1;
    }
    lb2_params P;
    lb2_default_params(&P);
    P.verbose = 1;
    int device = 0, opt, idx = 0;
    const char *save_binary = nullptr;      // write the parsed instance as a binary cache (readable in place of the .dat-s)
    const char *fname = argv[1];
    while ((opt = getopt_long(argc, argv, "", long_options, &idx)) != -1) {
        switch (opt) {
        case 1000: P.initRho = atof(optarg); break;       case 1001: P.rhoMax = atof(optarg); break;
        case 1002: P.rhoCellingALM = atof(optarg); break; case 1003: P.rhoCellingADMM = atof(optarg); break;
        case 1004: P.maxALMIter = atoll(optarg); break;   case 1005: P.maxADMMIter = atoll(optarg); break;
        case 1006: P.timesLogRank = atof(optarg); break;  case 1007: P.rhoFreq = atoll(optarg); break;
        case 1008: P.rhoFactor = atof(optarg); break;     case 1009: P.ALMRhoFactor = atof(optarg); break;
        case 1010: P.phase1Tol = atof(optarg); break;     case 1011: P.phase2Tol = atof(optarg); break;
        case 1012: P.timeSecLimit = atof(optarg); break;  case 1013: P.heuristicFactor = atof(optarg); break;
        case 1014: P.lbfgsListLength = atoll(optarg); break; case 1015: P.endTauTol = atof(optarg); break;
        case 1016: P.endALMSubTol = atof(optarg); break;  case 1017: P.l2Rescaling = atoi(optarg); break;
        case 1018: P.reoptLevel = atoll(optarg); break;   case 1019: P.dyrankLevel = atoll(optarg); break;
        case 1020: P.highAccMode = atoi(optarg); break;   case 2000: device = atoi(optarg); break;
        case 2001: P.verbose = 0; break;
        case 2002: save_binary = optarg; break;
        default: break;
        }
    }
    P.rhoCellingADMM = P.rhoMax * 200;      // main.c:236
    lb2_sdpa *F = nullptr;
    if (lb2_read_sdpa(fname, &F) != LB2_OK) { std::fprintf(stderr, "cannot read %s: %s\n", fname, lb2_sdpa_last_error()); return 2; }
    if (save_binary && lb2_sdpa_save(F, save_binary) != LB2_OK) { std::fprintf(stderr, "cannot write %s: %s\n", save_binary, lb2_sdpa_last_error()); return 2; }
    const lb2_int m = lb2_sdpa_info(F, 0, 0), nblk = lb2_sdpa_info(F, 1, 0);
    const lb2_int nlp = lb2_sdpa_info(F, 4, 0);
    std::printf("nConstrs = %lld, sdp nBlks = %lld, lp Cols = %lld\n", (long long)m, (long long)nblk, (long long)nlp);
    std::vector<lb2_int> dims((size_t)nblk);
    for (lb2_int k = 0; k < nblk; ++k) dims[(size_t)k] = lb2_sdpa_info(F, 2, k);
    std::vector<double> rhs((size_t)m);
    lb2_sdpa_get(F, -1, nullptr, nullptr, nullptr, rhs.data());
    lb2_solver *S = nullptr;
    if (lb2_create(&S, m, nblk, dims.data(), rhs.data(), device) != LB2_OK) fail("lb2_create");
    for (lb2_int k = 0; k < nblk; ++k) {
        const lb2_int nnz = lb2_sdpa_info(F, 3, k);
        std::vector<lb2_int> beg((size_t)m + 2), ix((size_t)nnz + 1);
        std::vector<double> el((size_t)nnz + 1);
        lb2_sdpa_get(F, k, beg.data(), ix.data(), el.data(), nullptr);
        if (lb2_set_cone_data(S, k, beg.data(), ix.data(), el.data()) != LB2_OK) fail("lb2_set_cone_data");
    }
    if (nlp > 0) {
        const lb2_int nnz = lb2_sdpa_info(F, 3, nblk);
        std::vector<lb2_int> beg((size_t)m + 2), ix((size_t)nnz + 1);
        std::vector<double> el((size_t)nnz + 1);
        lb2_sdpa_get(F, nblk, beg.data(), ix.data(), el.data(), nullptr);
        if (lb2_set_lp_data(S, nlp, beg.data(), ix.data(), el.data()) != LB2_OK) fail("lb2_set_lp_data");
    }
    lb2_sdpa_free(F);
    if (lb2_preprocess(S) != LB2_OK) fail("lb2_preprocess");
    if (lb2_determine_rank(S, P.timesLogRank) != LB2_OK) fail("lb2_determine_rank");
    if (lb2_init_vars(S, P.lbfgsListLength, P.initRho) != LB2_OK) fail("lb2_init_vars");
    lb2_result R;
    if (lb2_solve(S, &P, &R) != LB2_OK) fail("lb2_solve");
    static const char *why[] = {"the status is unknown", "`Official terminate criteria`", "`final terminate criteria`",
                                "`the maximum number of iterations`", "the time limit"};
    std::printf("-----------------------------------------------------------------------\n");
    std::printf("End Program due to reaching %s:\n", why[R.status]);
    std::printf("Objective function Value are:\n");
    std::printf("\t 1.Primal Objective:            : %10.6e\n", R.pObj);
    std::printf("\t 2.Dual Objective:              : %10.6e\n", R.dObj);
    std::printf("Dimacs Error are:\n");
    std::printf("\t 1.Constraint Violation(1)      : %10.6e\n", R.pInfeasL1);
    std::printf("\t 2.Dual Infeasibility(1)        : %10.6e\n", R.dInfeasL1);
    std::printf("\t 3.Primal Dual Gap              : %10.6e\n", R.pdGap);
    std::printf("\t 4.Primal Variable Semidefinite : %10.6e\n", 0.0);
    std::printf("\t 5.Constraint Violation(Inf)    : %10.6e\n", R.pInfeasInf);
    std::printf("\t 6.Dual Infeasibility(Inf)      : %10.6e\n", R.dInfeasInf);
    std::printf("-----------------------------------------------------------------------\n");
    std::printf("ALM inner iterations: %lld  ADMM iterations: %lld  CG iterations: %lld  final rank (cone 0): %lld\n",
                (long long)R.almInnerIter, (long long)R.admmIter, (long long)R.cgIter, (long long)R.finalRank0);
    std::printf("all_time: %f (ALM %f, ADMM %f)  kernel launches: %lld\n", R.solveSeconds, R.almSeconds, R.admmSeconds,
                (long long)R.kernelLaunches);
    lb2_destroy(S);
    return 0;
}
