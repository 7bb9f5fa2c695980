// Host-side driver pieces of the reference restated in C++ (row f1 of SURVEY.md section 8):
// data-file reader, sort, report formatting.  One-time host work around the GPU hot path.
#pragma once
#include <string>
#include <vector>

namespace fmc_host {

// gfortran list-directed output of a REAL(8): WRITE(uo,*) value  (isolated here: the exact
// layout is compiler-defined and could not be verified without a Fortran compiler)
std::string list_directed_real(double v);

// utils.f90:40-94: print text with line breaks only at spaces, lines of <= linelength chars,
// each written list-directed (so with a leading blank).
std::string wordwrap(const std::string& text, int linelength = 79);

// utils.f90:128-218: "0.123546(7)" notation.
std::string write_mean(double av, double std_err_in_mean, int err_prec = 1);

// utils.f90:239-335 (Numerical-Recipes style incomplete gamma function Q(a,x)).
double gammq(double a, double x);

// utils.f90:97-105: does the first word of the line read as a number (Fortran list-directed)?
bool isdataline(const std::string& line);

struct DataSet {
  std::vector<double> x, y, err;
  bool was_out_of_order = false;
  bool has_equal_x = false;
  double equal_x_value = 0.0;
};

// bootstrap.f90:54-152 + 155-194: read three columns, reject err<=0, sort by x (stable).
// Returns "" on success or the reference's error message.
std::string read_data(const std::string& path, DataSet& out);

}  // namespace fmc_host
