// fitter_mc -- C++ stand-in for the reference's driver program (fitter_mc.f90 + the host parts
// of bootstrap.f90), for hosts without a Fortran compiler: `./fitter_mc datafile.dat` prints the
// reference's transcript with the Monte-Carlo loop running on the GPU
// through the C ABI of libfmc_b200.so.  With gfortran available the reference's own driver is
// used instead, bound as INTEGRATION.md shows.
//
// Environment (extensions; defaults reproduce the reference): FMC_NRAND (nrand_points, default
// 100000, bootstrap.f90:9), FMC_SEED (31415, utils.f90:7), FMC_NGPUS (1), FMC_MODEL (0),
// FMC_JACOBIAN (analytic = what the reference's README:35-42 calls replacing nl2sno by nl2sol with a
// madj: exact Jacobian rows, from the plug-in's model_grad or by dual numbers; default: forward
// differences, the reference's path),
// FMC_HISTOGRAM (1 = write histogram_<prop>.dat, bootstrap.f90:263), FMC_HISTOGRAM_BINS (> 0:
// bin the properties on the GPU instead of dumping one text line per resample -- the form that
// scales to 1e8 resamples; histogram_<prop>.dat then holds "bin centre, count" lines).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/fmc_b200.h"
#include "report.hpp"

using namespace fmc_host;

static void errstop(const std::string& sub, const std::string& message) {  // utils.f90:26-37
  std::fprintf(stderr, "\n ERROR in subroutine %s.\n\n%s\n", sub.c_str(), wordwrap(message).c_str());
  std::fflush(stderr);
  std::exit(1);
}

struct Model {
  int id = 0, nparam = 0, nprop = 0;
  char pname[FMC_MAX_PARAM][3];
  char qname[FMC_MAX_PARAM][3];
  char name[32];
  std::vector<double> init;
};

static std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(' '), b = s.find_last_not_of(' ');
  return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}

// display_params, bootstrap.f90:410-462
static void display_params(const Model& m, const DataSet& d, const double* params, const double* err_params,
                           const double* props, const double* err_props) {
  for (int j = 0; j < m.nparam; ++j) {
    if (err_params) {
      if (err_params[j] > 0.0) {
        std::printf("   Parameter %s = %s +/- %s\n", m.pname[j], list_directed_real(params[j]).c_str(),
                    list_directed_real(err_params[j]).c_str());
        std::printf("                = %s\n", write_mean(params[j], err_params[j]).c_str());
      } else {
        std::printf(" Parameter %s = %s\n", m.pname[j], list_directed_real(params[j]).c_str());
      }
    } else {
      std::printf("   Parameter %s = %s\n", m.pname[j], list_directed_real(params[j]).c_str());
    }
  }
  if (m.nprop > 0 && props) {
    for (int j = 0; j < m.nprop; ++j) {
      if (err_props) {
        if (err_props[j] > 0.0) {
          std::printf("   Property  %s = %s +/- %s\n", m.qname[j], list_directed_real(props[j]).c_str(),
                      list_directed_real(err_props[j]).c_str());
          std::printf("                = %s\n", write_mean(props[j], err_props[j]).c_str());
        } else {
          std::printf(" Property  %s = %s\n", m.qname[j], list_directed_real(props[j]).c_str());
        }
      } else {
        std::printf("   Property  %s = %s\n", m.qname[j], list_directed_real(props[j]).c_str());
      }
    }
  }
  double chi2 = 0.0;  // bootstrap.f90:452-458
  const int n = (int)d.x.size();
  for (int i = 0; i < n; ++i) {
    const double r = (d.y[i] - fmc_model_y(m.id, params, d.x[i])) * (1.0 / d.err[i]);
    chi2 += r * r;
  }
  std::printf("   Reduced chi^2 function = %s\n", list_directed_real(chi2 / (double)(n - m.nparam)).c_str());
  if (n > m.nparam)
    std::printf("   Probability chi^2 is this bad = %s\n",
                list_directed_real(gammq((double)(n - m.nparam) * 0.5, chi2 * 0.5)).c_str());
}

int main(int argc, char** argv) {
  std::printf("\n FITTER_MC\n =========\n\n");  // fitter_mc.f90:21-30 (non-OpenMP build)

  // get_file, bootstrap.f90:30-51
  std::string in_file;
  if (argc > 1) {
    in_file = argv[1];
  } else {
    for (;;) {
      std::printf(" Please enter name of data file.\n");
      if (std::cin >> in_file) break;
      if (std::cin.eof()) errstop("GET_FILE", "No file name given.");
      std::printf(" Please try again.\n");
      std::cin.clear();
    }
    std::printf("\n");
  }
  in_file = trim(in_file);
  std::fputs(wordwrap("Reading data from " + in_file + ".").c_str(), stdout);

  DataSet d;
  const std::string e = read_data(in_file, d);
  if (!e.empty()) {
    if (e.rfind("Found a non-positive", 0) == 0 || e.rfind("Problem reading", 0) == 0) {
      std::printf(" %s\n", e.c_str());
      errstop("GET_DATA", "Halting.");
    }
    errstop("GET_DATA", e);
  }
  const int ndata = (int)d.x.size();
  std::printf(" Number of data lines: %d\n\n", ndata);
  if (d.was_out_of_order) std::printf(" Warning: raw x data are not in ascending order.\n\n");
  if (d.has_equal_x)
    std::printf(" Warning: two x data are equal.\n Value: %s\n\n", list_directed_real(d.equal_x_value).c_str());

  // initialise_model, bootstrap.f90:365-387
  Model m;
  m.id = std::getenv("FMC_MODEL") ? std::atoi(std::getenv("FMC_MODEL")) : 0;
  m.init.resize(FMC_MAX_PARAM);
  if (fmc_initialise_model(m.id, &m.nparam, &m.nprop, m.init.data(), m.pname, m.qname, m.name) != FMC_OK)
    errstop("INITIALISE_MODEL", "Unknown model id.");
  std::printf(" Model used is: %s\n", m.name);
  std::printf(" Number of parameters: %d\n\n", m.nparam);
  if (ndata < m.nparam) errstop("MONTE_CARLO_FIT_DATA", "Fewer data points than parameters.");

  if (std::getenv("FMC_JACOBIAN") && std::string(std::getenv("FMC_JACOBIAN")) == "analytic") {
    if (!fmc_model_has_analytic_jacobian(m.id))
      errstop("MONTE_CARLO_FIT_DATA", "This model has no derivative (model_y is not generic and there is no model_grad).");
    fmc_set_jacobian_mode(nullptr, FMC_JACOBIAN_ANALYTIC);
  }

  std::vector<double> params(m.init.begin(), m.init.begin() + m.nparam), props(m.nprop > 0 ? m.nprop : 1);
  std::printf(" Initial parameter set:\n");
  display_params(m, d, params.data(), nullptr, nullptr, nullptr);
  std::printf("\n");

  // initial fit with non-offset data, bootstrap.f90:288-295
  int rc2[2];
  int rc = fmc_fit_once(m.id, ndata, d.x.data(), d.y.data(), d.err.data(), params.data(), rc2);
  if (rc != FMC_OK) errstop("MONTE_CARLO_FIT_DATA", std::string(fmc_strerror(rc)) + " " + fmc_last_cuda_error());
  fmc_derived_properties(m.id, params.data(), props.data());
  std::printf(" Fitted parameter set:\n");
  display_params(m, d, params.data(), nullptr, props.data(), nullptr);
  std::printf("\n");

  // the Monte Carlo loop, bootstrap.f90:297-330: one call
  const long long nrand = std::getenv("FMC_NRAND") ? std::atoll(std::getenv("FMC_NRAND")) : 100000LL;
  const unsigned long long seed = std::getenv("FMC_SEED") ? std::strtoull(std::getenv("FMC_SEED"), nullptr, 10) : FMC_DEFAULT_SEED;
  const int ngpus = std::getenv("FMC_NGPUS") ? std::atoi(std::getenv("FMC_NGPUS")) : 1;
  const bool make_histogram = !(std::getenv("FMC_HISTOGRAM") && std::atoi(std::getenv("FMC_HISTOGRAM")) == 0);
  std::vector<double> sp(m.nparam), sp2(m.nparam), sq(props.size()), sq2(props.size());
  std::vector<double> props_all;
  const int nbins = std::getenv("FMC_HISTOGRAM_BINS") ? std::atoi(std::getenv("FMC_HISTOGRAM_BINS")) : 0;
  const bool binned = make_histogram && nbins > 0 && m.nprop > 0;
  const int nquant = m.nparam + m.nprop;
  std::vector<double> blo(nquant), bhi(nquant);
  std::vector<unsigned long long> counts;
  if (binned) {
    // range of every quantity from a pilot of a few thousand resamples (ids from the far end of
    // the id space, so the run proper does not repeat them): mean -/+ 8 standard deviations
    const long long npilot = 4096;
    std::vector<double> pp((size_t)npilot * m.nparam), pq((size_t)npilot * m.nprop);
    rc = fmc_bootstrap_run(m.id, ndata, d.x.data(), d.y.data(), d.err.data(), params.data(), npilot, seed,
                           (1LL << 62) - npilot, 1, nullptr, sp.data(), sp2.data(), sq.data(), sq2.data(), pp.data(),
                           pq.data(), nullptr, nullptr);
    if (rc != FMC_OK) errstop("MONTE_CARLO_FIT_DATA", std::string(fmc_strerror(rc)) + " " + fmc_last_cuda_error());
    for (int j = 0; j < nquant; ++j) {
      const double* v = j < m.nparam ? pp.data() + j : pq.data() + (j - m.nparam);
      const int stride = j < m.nparam ? m.nparam : m.nprop;
      double mean = 0.0, var = 0.0;
      long long cnt = 0;
      for (long long i = 0; i < npilot; ++i)
        if (std::isfinite(v[i * stride])) {
          mean += v[i * stride];
          ++cnt;
        }
      mean /= (double)(cnt > 0 ? cnt : 1);
      for (long long i = 0; i < npilot; ++i)
        if (std::isfinite(v[i * stride])) var += (v[i * stride] - mean) * (v[i * stride] - mean);
      double sd = std::sqrt(var / (double)(cnt > 1 ? cnt - 1 : 1));
      if (!(sd > 0.0)) sd = 0.0625;
      blo[j] = mean - 8.0 * sd;
      bhi[j] = mean + 8.0 * sd;
    }
    counts.resize((size_t)nquant * (nbins + 2));
    rc = fmc_bootstrap_run_binned(m.id, ndata, d.x.data(), d.y.data(), d.err.data(), params.data(), nrand, seed, 0,
                                  ngpus, sp.data(), sp2.data(), sq.data(), sq2.data(), nullptr, nullptr, nbins,
                                  blo.data(), bhi.data(), counts.data());
  } else {
    if (make_histogram && m.nprop > 0) props_all.resize((size_t)nrand * m.nprop);
    rc = fmc_bootstrap_run(m.id, ndata, d.x.data(), d.y.data(), d.err.data(), params.data(), nrand, seed, 0, ngpus,
                           nullptr, sp.data(), sp2.data(), sq.data(), sq2.data(), nullptr,
                           props_all.empty() ? nullptr : props_all.data(), nullptr, nullptr);
  }
  if (rc != FMC_OK) errstop("MONTE_CARLO_FIT_DATA", std::string(fmc_strerror(rc)) + " " + fmc_last_cuda_error());
  if (binned) {
    for (int j = 0; j < m.nprop; ++j) {
      const std::string fn = "histogram_" + trim(m.qname[j]) + ".dat";
      FILE* f = std::fopen(fn.c_str(), "w");
      if (!f) errstop("MONTE_CARLO_FIT_DATA", "Error opening " + fn + ".");
      const int q = m.nparam + j;
      const unsigned long long* c = counts.data() + (size_t)q * (nbins + 2);
      const double width = (bhi[q] - blo[q]) / (double)nbins;
      std::fprintf(f, "# binned on the GPU: %d bins on [%.17g, %.17g); below: %llu; above: %llu\n", nbins, blo[q],
                   bhi[q], c[0], c[nbins + 1]);
      for (int b = 1; b <= nbins; ++b)
        std::fprintf(f, "%s %llu\n", list_directed_real(blo[q] + width * ((double)b - 0.5)).c_str(), c[b]);
      std::fclose(f);
    }
  } else if (make_histogram) {
    for (int j = 0; j < m.nprop; ++j) {
      const std::string fn = "histogram_" + trim(m.qname[j]) + ".dat";
      FILE* f = std::fopen(fn.c_str(), "w");
      if (!f) errstop("MONTE_CARLO_FIT_DATA", "Error opening " + fn + ".");
      for (long long i = 0; i < nrand; ++i)
        std::fprintf(f, "%s\n", list_directed_real(props_all[(size_t)i * m.nprop + j]).c_str());
      std::fclose(f);
    }
  }
  // bootstrap.f90:328-341
  std::vector<double> pm(m.nparam), pe(m.nparam), qm(props.size()), qe(props.size());
  fmc_finalise_moments(m.nparam, nrand, sp.data(), sp2.data(), pm.data(), pe.data());
  fmc_finalise_moments(m.nprop, nrand, sq.data(), sq2.data(), qm.data(), qe.data());
  std::printf(" Monte Carlo fitted parameter set (using %lld random points):\n", nrand);
  display_params(m, d, pm.data(), pe.data(), qm.data(), qe.data());
  std::printf("\n");
  if (make_histogram) {
    for (int j = 0; j < m.nprop; ++j)
      std::printf("   Histogram of %s written to histogram_%s.dat.\n", trim(m.qname[j]).c_str(), trim(m.qname[j]).c_str());
    std::printf("\n");
  }
  if (binned) {  // extension: central 68% / 95% intervals straight from the binned distributions
    const double probs[5] = {0.025, 0.16, 0.5, 0.84, 0.975};
    for (int j = 0; j < m.nprop; ++j) {
      const int q = m.nparam + j;
      double qs[5];
      fmc_histogram_quantiles(counts.data() + (size_t)q * (nbins + 2), nbins, blo[q], bhi[q], probs, 5, qs);
      std::printf("   Quantiles of %s (2.5, 16, 50, 84, 97.5 %%):\n  ", m.qname[j]);
      for (int k = 0; k < 5; ++k) std::printf(" %.10g", qs[k]);
      std::printf("\n");
    }
    std::printf("\n");
  }
  std::printf(" Program finished.\n\n");
  return 0;
}
