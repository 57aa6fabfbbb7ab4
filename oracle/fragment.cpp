// ORACLE — TEST INFRASTRUCTURE ONLY.  See fragment.hpp for the reference lines restated here.
#include "fragment.hpp"
#include <cassert>
#include <cmath>

namespace oracle {

// Wichura's algorithm AS 241 (Appl. Statist. 37 (1988) 477-484), routine PPND16: three rational
// minimax approximations, central (|q| <= 0.425), intermediate (r <= 5) and far tail.  The reference
// takes it from R's qnorm.c (statistics.cpp:47-113).
double Phi_inv(double p) {
    assert(0.0 < p && p < 1.0);
    static const double a[8] = {3.387132872796366608, 133.14166789178437745, 1971.5909503065514427, 13731.693765509461125,
                                45921.953931549871457, 67265.770927008700853, 33430.575583588128105, 2509.0809287301226727};
    static const double b[8] = {1.0, 42.313330701600911252, 687.1870074920579083, 5394.1960214247511077,
                                21213.794301586595867, 39307.89580009271061, 28729.085735721942674, 5226.495278852854561};
    static const double c[8] = {1.42343711074968357734, 4.6303378461565452959, 5.7694972214606914055, 3.64784832476320460504,
                                1.27045825245236838258, 0.24178072517745061177, 0.0227238449892691845833, 7.7454501427834140764e-4};
    static const double d[8] = {1.0, 2.05319162663775882187, 1.6763848301838038494, 0.68976733498510000455,
                                0.14810397642748007459, 0.0151986665636164571966, 5.475938084995344946e-4, 1.05075007164441684324e-9};
    static const double e[8] = {6.6579046435011037772, 5.4637849111641143699, 1.7848265399172913358, 0.29656057182850489123,
                                0.026532189526576123093, 0.0012426609473880784386, 2.71155556874348757815e-5, 2.01033439929228813265e-7};
    static const double f[8] = {1.0, 0.59983220655588793769, 0.13692988092273580531, 0.0148753612908506148525,
                                7.868691311456132591e-4, 1.8463183175100546818e-5, 1.4215117583164458887e-7, 2.04426310338993978564e-15};
    auto horner = [](const double* k, double r) { double v = k[7]; for (int i = 6; i >= 0; i--) v = v * r + k[i]; return v; };
    const double q = p - 0.5;
    if (std::fabs(q) <= 0.425) {
        const double r = 0.180625 - q * q;
        return q * horner(a, r) / horner(b, r);
    }
    double r = q > 0 ? 1.0 - p : p;
    r = std::sqrt(-std::log(r));
    double val;
    if (r <= 5.0) { r -= 1.6; val = horner(c, r) / horner(d, r); }
    else { r -= 5.0; val = horner(e, r) / horner(f, r); }
    return q < 0.0 ? -val : val;
}

static double normal_pdf(double x, double m, double s) {
    const double z = (x - m) / s;
    return 0.3989422804014327 / s * std::exp(-0.5 * z * z);
}

FragmentLengthDistribution::FragmentLengthDistribution(size_t maximum_sample_size_, size_t reestimation_frequency_, double fraction)
    : robust_estimation_fraction(fraction), maximum_sample_size(maximum_sample_size_), reestimation_frequency(reestimation_frequency_) {
    assert(0.0 < fraction && fraction < 1.0);
}

void FragmentLengthDistribution::force_parameters(double mean, double stddev) { mu = mean; sigma = stddev; is_fixed = true; }

void FragmentLengthDistribution::register_fragment_length(int64_t length) {
    if (is_fixed) return;
    lengths.insert((double)length);
    if (lengths.size() == maximum_sample_size) { estimate_distribution(); is_fixed = true; }
    else if (lengths.size() % reestimation_frequency == 0) estimate_distribution();
}

void FragmentLengthDistribution::estimate_distribution() {
    // drop both tails, then method of moments for the truncated normal
    const size_t to_skip = (size_t)(lengths.size() * (1.0 - robust_estimation_fraction) * 0.5);
    auto begin = lengths.begin(); auto end = lengths.end();
    for (size_t i = 0; i < to_skip; i++) { ++begin; --end; }
    double count = 0.0, sum = 0.0, sum_of_sqs = 0.0;
    for (auto it = begin; it != end; ++it) { count += 1.0; sum += *it; sum_of_sqs += (*it) * (*it); }
    mu = sum / count;
    const double raw_var = sum_of_sqs / count - mu * mu;
    const double a = Phi_inv(1.0 - 0.5 * (1.0 - robust_estimation_fraction));
    sigma = std::sqrt(raw_var / (1.0 - 2.0 * a * normal_pdf(a, 0.0, 1.0)));
}

} // namespace oracle

extern "C" double oracle_phi_inv(double p) { return oracle::Phi_inv(p); }

// Register `n` lengths in order; reports the state afterwards.
extern "C" void oracle_fragment_estimate(const int64_t* lengths, uint64_t n, uint64_t maximum_sample_size, uint64_t reestimation_frequency,
                                         double robust_estimation_fraction, double* mean, double* stdev, int* finalized, uint64_t* sample_size) {
    oracle::FragmentLengthDistribution dist(maximum_sample_size, reestimation_frequency, robust_estimation_fraction);
    for (uint64_t i = 0; i < n; i++) dist.register_fragment_length(lengths[i]);
    *mean = dist.mean(); *stdev = dist.std_dev(); *finalized = dist.is_finalized() ? 1 : 0; *sample_size = dist.curr_sample_size();
}
