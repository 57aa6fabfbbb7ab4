// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product).
// FragmentLengthDistribution, restated from the reference:
//   class                       mapper.hpp:83-139
//   register_fragment_length    mapper.cpp:5256-5278
//   estimate_distribution       mapper.cpp:5280-5305
//   force_parameters            mapper.cpp:5250-5254
//   Phi_inv (AS 241, PPND16)    statistics.cpp:49-113;  normal_pdf statistics.hpp:181-188
// Pinned by the reference's own fixture (unittest/minimizer_mapper.cpp:37-108, committed as
// tests/golden/fragment_lengths.json) and by scipy's norm.ppf for Phi_inv.
#pragma once
#include <cstddef>
#include <cstdint>
#include <set>

namespace oracle {

double Phi_inv(double p);

class FragmentLengthDistribution {
public:
    FragmentLengthDistribution(size_t maximum_sample_size, size_t reestimation_frequency, double robust_estimation_fraction);
    void force_parameters(double mean, double stddev);
    void register_fragment_length(int64_t length);
    double mean() const { return mu; }
    double std_dev() const { return sigma; }
    bool is_finalized() const { return is_fixed; }
    size_t max_sample_size() const { return maximum_sample_size; }
    size_t curr_sample_size() const { return lengths.size(); }
private:
    std::multiset<double> lengths;
    bool is_fixed = false;
    double robust_estimation_fraction;
    size_t maximum_sample_size, reestimation_frequency;
    double mu = 0.0, sigma = 1.0;
    void estimate_distribution();
};

} // namespace oracle
