#ifndef _DYMU_CORA_HPP_
#define _DYMU_CORA_HPP_

// DyMuCoRa.hpp -- running statistics behind the "cost ratio after traverse" (CoRa) methods.
//
// Host-side restatement of the sample bookkeeping of the reference's src/DyMu.hpp:110-394
// (costCriteria, segmentedTerrain).  These are a few dozen scalars per terrain class and stay
// on the CPU; what this build moves to the GPU is the consequence of a changed look-up table
// (DyMuPathPlanner::recomputeCostMap: cost-map rebuild from the resident DEM + re-solve).
//
// Type and member names are the reference's so that callers that inspect
// terrain_vector[t].criteria_info[c].mean keep compiling.  The arithmetic keeps the reference's
// operation order (the pooled deviation is not the textbook one: the cross term uses the old
// mean, and merging with an empty accumulator runs with n-1 = -1), because updateCost() is
// compared value for value against the compiled reference in tests/test_cora_vs_reference.py.

#include <cmath>
#include <cstdlib>
#include <iostream>
#include <vector>

#include <base-logging/Logging.hpp>

namespace PathPlanning_lib
{
// One accumulator: count, mean, sample standard deviation.  reference: H.hpp:110-204
struct costCriteria
{
    int num_samples;
    double mean;
    double std_deviation;
    bool empty;

    costCriteria() { erase(); }
    costCriteria(int num_samples_, double mean_, double std_deviation_)
        : num_samples(num_samples_), mean(mean_), std_deviation(std_deviation_), empty(false)
    {
    }

    void erase()
    {
        num_samples = 0;
        mean = 0;
        std_deviation = 0;
        empty = true;
    }

    // fold a batch of raw samples in (H.hpp:131-166)
    void addData(std::vector<double> new_samples)
    {
        const int n = (int)new_samples.size();
        if (n == 0) return;
        double sum = 0;
        for (int k = 0; k < n; ++k) sum += new_samples[k];
        const double merged_mean = (mean * num_samples + sum) / (num_samples + n);
        if (num_samples + n - 2 > 0)
        {
            double cross = 0;
            for (int k = 0; k < n; ++k)
            {
                const double x = new_samples[k];
                cross += empty ? std::pow(x - merged_mean, 2) : (x - mean) * (x - merged_mean);
            }
            std_deviation =
                std::sqrt((std::pow(std_deviation, 2) * (num_samples - 1) + cross) / (num_samples + n - 2));
        }
        else
            LOG_ERROR_S << "ERROR: not enough samples to obtain standard deviation.";
        num_samples += n;
        mean = merged_mean;
        empty = false;
    }

    // fold an already summarised group in (H.hpp:168-184)
    void addData(int num_samples_, double mean_, double std_deviation_)
    {
        if (num_samples_ == 0) return;
        const double merged_mean = (mean * num_samples + mean_ * num_samples_) / (num_samples + num_samples_);
        std_deviation = std::sqrt((std::pow(std_deviation, 2) * (num_samples - 1)
                                   + std::pow(std_deviation_, 2) * (num_samples_ - 1))
                                  / (num_samples + num_samples_ - 2));
        num_samples += num_samples_;
        mean = merged_mean;
        empty = false;
    }

    // one more sample (H.hpp:186-198).  As in the reference the cross term is only defined
    // for a non-empty accumulator; an empty one contributes 0 here instead of an
    // uninitialised value.
    void addData(double new_sample)
    {
        const double merged_mean = (mean * num_samples + new_sample) / (num_samples + 1);
        const double cross = empty ? 0.0 : (new_sample - mean) * (new_sample - merged_mean);
        std_deviation = std::sqrt((std::pow(std_deviation, 2) * (num_samples - 1) + cross) / (num_samples + 1 - 2));
        num_samples += 1;
        mean = merged_mean;
        empty = false;
    }
};

// Everything learnt about one terrain class.  reference: H.hpp:206-394
struct segmentedTerrain
{
    double cost;
    double slope_ratio;  // cost increase per degree of slope
    std::vector<costCriteria> criteria_info, traverse_info, rejected_info;
    std::vector<std::vector<double>> data_samples;
    bool traversed;

    segmentedTerrain() : cost(1), slope_ratio(1), traversed(false) {}
    segmentedTerrain(double cost_, double slope_ratio_) : cost(cost_), slope_ratio(slope_ratio_), traversed(false) {}
    segmentedTerrain(std::vector<costCriteria> criteria_info_)
        : criteria_info(criteria_info_),
          traverse_info(criteria_info_.size()),
          rejected_info(criteria_info_.size()),
          data_samples(criteria_info_.size()),
          traversed(true)
    {
    }

    enum
    {
        kEnoughSamples = 29,  // "> 29" marks a criterion as established
        kBootstrapBatch = 2,  // "> 2" raw samples are folded while bootstrapping
        kTraverseBatch = 9    // "> 9" raw samples form one traverse group afterwards
    };

    // Decide what happens to the raw samples gathered since the last call (H.hpp:234-309).
    void dataAnalysis()
    {
        const size_t nc = criteria_info.size();
        if (!traversed)
        {
            for (size_t c = 0; c < nc; ++c)
            {
                if (data_samples[c].size() > kBootstrapBatch) absorb(criteria_info[c], c);
                if (criteria_info[c].num_samples > kEnoughSamples)
                {
                    traversed = true;
                    std::cout << "\033[1;32mNow we have gathered enough info of the current terrain.\033[0m"
                              << std::endl;
                }
            }
            return;
        }
        for (size_t c = 0; c < nc; ++c)
        {
            costCriteria& kept = criteria_info[c];
            if (kept.num_samples <= kEnoughSamples)
            {
                absorb(kept, c);
                continue;
            }
            if (data_samples[c].size() > kTraverseBatch)
            {
                costCriteria& group = traverse_info[c];
                group.addData(data_samples[c]);
                if (FTest((int)c)) kept.addData(group.num_samples, group.mean, group.std_deviation);
                data_samples[c].clear();
                group.erase();
            }
            costCriteria& rejected = rejected_info[c];
            if (rejected.num_samples > kEnoughSamples)
            {
                if (TTest((int)c))
                    kept.addData(rejected.num_samples, rejected.mean, rejected.std_deviation);
                else if (rejected.num_samples >= kept.num_samples && rejected.std_deviation < kept.std_deviation)
                {
                    LOG_WARN_S << "WARNING: [Criteria " << c + 1
                               << "] rejected samples outnumber and outclass the kept ones: swapping them";
                    // swap kept <-> rejected by re-accumulating, exactly as the reference does
                    // (merging into an emptied accumulator rescales the deviation by
                    // sqrt((n-1)/(n-2)): kept)
                    costCriteria& tmp = traverse_info[c];
                    tmp.erase();
                    tmp.addData(kept.num_samples, kept.mean, kept.std_deviation);
                    kept.erase();
                    kept.addData(rejected.num_samples, rejected.mean, rejected.std_deviation);
                    rejected.erase();
                    rejected.addData(tmp.num_samples, tmp.mean, tmp.std_deviation);
                    tmp.erase();
                }
            }
        }
    }

    // kept vs rejected, both large groups (H.hpp:311-327).  The reference calls the unqualified
    // abs() on a double inside the header, i.e. before its .cpp files include <math.h>; unless an
    // earlier header already exported std::abs into the global namespace, lookup finds only
    // ::abs(int) and the difference of the means is truncated to an integer.  That is what the
    // compiled reference (oracle/_ref) does and what is reproduced here; build with
    // -DDYMU_CORA_FLOAT_ABS for the floating-point reading.
    static double meanGap(double d)
    {
#ifdef DYMU_CORA_FLOAT_ABS
        return std::fabs(d);
#else
        return (double)std::abs((int)d);
#endif
    }
    bool TTest(int i)
    {
        const costCriteria &a = criteria_info[i], &b = rejected_info[i];
        const double n1 = a.num_samples, n2 = b.num_samples;
        const double t =
            meanGap(a.mean - b.mean) / std::sqrt(std::pow(a.std_deviation, 2) / n1 + std::pow(b.std_deviation, 2) / n2);
        return t < 2.00;
    }

    // variance ratio picks the mean test (H.hpp:329-343)
    bool FTest(int i)
    {
        const double F = std::pow(traverse_info[i].std_deviation, 2) / std::pow(criteria_info[i].std_deviation, 2);
        return F < 2.05 ? studentTTest(i) : cochranTTest(i);
    }

    // similar deviations: pooled one-sided t (H.hpp:345-369); a rejected group is remembered
    bool studentTTest(int i)
    {
        const costCriteria &a = criteria_info[i], &b = traverse_info[i];
        const double n1 = a.num_samples, n2 = b.num_samples;
        const double sp = std::sqrt(((n1 - 1) * std::pow(a.std_deviation, 2) + (n2 - 1) * std::pow(b.std_deviation, 2))
                                    / (n1 + n2 - 2));
        const double t = std::sqrt(n1 * n2 / (n1 + n2)) * (a.mean - b.mean) / sp;
        if (t < 2.02) return true;
        LOG_WARN_S << "WARNING: [Criteria " << i + 1 << "] Sample rejected after Student T test.";
        rejected_info[i].addData(b.num_samples, b.mean, b.std_deviation);
        return false;
    }

    // different deviations: Cochran's approximation (H.hpp:371-393); a rejected group is dropped
    bool cochranTTest(int i)
    {
        const costCriteria &a = criteria_info[i], &b = traverse_info[i];
        const double n1 = a.num_samples, n2 = b.num_samples;
        const double q1 = std::pow(a.std_deviation, 2), q2 = std::pow(b.std_deviation, 2);
        const double tcal = (a.mean - b.mean) / std::sqrt(q1 / n1 + q2 / n2);
        const double ttab = (2.02 * q1 / n1 + 2.22 * q2 / n2) / (q1 / n1 + q2 / n2);
        if (tcal < ttab) return true;
        LOG_WARN_S << "WARNING: [Criteria " << i + 1 << "] Sample rejected after Cochran T test.";
        return false;
    }

  private:
    void absorb(costCriteria& into, size_t c)
    {
        into.addData(data_samples[c]);
        data_samples[c].clear();
    }
};

}  // namespace PathPlanning_lib

#endif  // _DYMU_CORA_HPP_
