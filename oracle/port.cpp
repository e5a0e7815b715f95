/* port.cpp -- CPU restatement of the reference hot path (plain C++17, no
 * reference headers).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h): only tests/, smoke() and the
 * cpu_baseline leg of bench.py may load the library built from this file.
 * Parity status: PINNED against the reference itself -- tests/test_oracle.py
 * compares every function here with oracle/_ref/libcsm_ref.so (the unmodified
 * reference sources compiled in this container) and with the golden vectors
 * in tests/golden/ that were generated from it (the reference ships no tests
 * or known-answer vectors of its own).
 *
 * Every function cites the reference file:line it restates (paths relative
 * to /root/reference). Arithmetic is done in the same order and precision as
 * the reference (double, no FMA contraction: built with -ffp-contract=off).
 */

#include "oracle_api.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <queue>
#include <thread>
#include <vector>

namespace {

/* ---- grid storage -------------------------------------------------------
 * grid_map_new/grid_map.hpp:27-264. The reference keeps 16x16 blocks that are
 * allocated on first write; a lookup outside the map or inside a block that
 * was never written returns the caller's default (grid_map.cpp:385-397,
 * 424-436). Here the grid is dense; a block counts as allocated iff it holds
 * a non-zero cell, which is how orc_grid_create fills the reference map. */
constexpr int kLog2Block = 4;
constexpr int kBlock = 1 << kLog2Block;

struct Grid
{
    int rows = 0, cols = 0;
    double res = 0.0, offX = 0.0, offY = 0.0;
    std::vector<uint16_t> cells;
    std::vector<uint8_t> blockAllocated;   /* (rows/16) x (cols/16) */
    bool allBlocks = false;                /* precomputed maps: every block */

    inline bool Inside(int row, int col) const
    { return row >= 0 && row < rows && col >= 0 && col < cols; }

    inline bool Allocated(int row, int col) const
    {
        if (!Inside(row, col))
            return false;
        if (allBlocks)
            return true;
        return blockAllocated[(row >> kLog2Block) * (cols >> kLog2Block) +
                              (col >> kLog2Block)] != 0;
    }

    /* GridMap::ValueOr, grid_map.cpp:385-397 */
    inline uint16_t ValueOr(int row, int col, uint16_t dflt) const
    {
        if (!Allocated(row, col))
            return dflt;
        return cells[static_cast<size_t>(row) * cols + col];
    }
};

/* ---- value <-> probability ----------------------------------------------
 * grid_binary_bayes.hpp:163-176 (constants), grid_values.hpp:26-36 (formula),
 * grid_values.cpp:11-46 (lookup table, entry 0 = unknown = 0.0). */
constexpr uint16_t kValueMin = 1;
constexpr uint16_t kValueMax = 65535;
constexpr double kProbMin = 1e-3;
constexpr double kProbMax = 1.0 - 1e-3;

struct ProbabilityTable
{
    std::vector<double> lut;
    ProbabilityTable() : lut(65536, 0.0)
    {
        /* The reference table has 65535 entries, so value 65535 indexes one
         * past its end. Entry 65535 is defined here by the
         * same formula; parity inputs are capped at 65534. */
        for (int v = 1; v <= 65535; ++v)
            lut[v] = kProbMin + (kProbMax - kProbMin) *
                     static_cast<double>(v - kValueMin) /
                     static_cast<double>(kValueMax - kValueMin);
    }
};

const ProbabilityTable gProb;

/* GridMap::ProbabilityOr, grid_map.cpp:424-436 */
inline double ProbabilityOr(const Grid& g, int row, int col, double dflt)
{
    if (!g.Allocated(row, col))
        return dflt;
    return gProb.lut[g.cells[static_cast<size_t>(row) * g.cols + col]];
}

/* ---- poses --------------------------------------------------------------- */
struct Pose { double x, y, t; };

/* pose.hpp:154-166 */
inline Pose Compound(const Pose& start, const Pose& diff)
{
    const double s = std::sin(start.t);
    const double c = std::cos(start.t);
    return Pose { c * diff.x - s * diff.y + start.x,
                  s * diff.x + c * diff.y + start.y,
                  start.t + diff.t };
}

/* pose.hpp:183-198 */
inline Pose InverseCompound(const Pose& start, const Pose& end)
{
    const double s = std::sin(start.t);
    const double c = std::cos(start.t);
    const double dx = end.x - start.x;
    const double dy = end.y - start.y;
    return Pose { c * dx + s * dy, -s * dx + c * dy, end.t - start.t };
}

/* pose.hpp:211-224 */
inline Pose MoveBackward(const Pose& end, const Pose& diff)
{
    const double theta = end.t - diff.t;
    const double s = std::sin(theta);
    const double c = std::cos(theta);
    return Pose { end.x - c * diff.x + s * diff.y,
                  end.y - s * diff.x - c * diff.y, theta };
}

/* ---- scans --------------------------------------------------------------- */
struct Scan
{
    std::vector<double> angles, ranges;
    Pose rel { 0.0, 0.0, 0.0 };
    size_t N() const { return ranges.size(); }
};

/* ScanData::HitPoint, sensor/sensor_data.hpp:190-203 */
inline void HitPoint(const Scan& scan, const Pose& sensor, size_t i,
                     double& hx, double& hy)
{
    const double c = std::cos(sensor.t + scan.angles[i]);
    const double s = std::sin(sensor.t + scan.angles[i]);
    hx = sensor.x + scan.ranges[i] * c;
    hy = sensor.y + scan.ranges[i] * s;
}

/* GridMapGeometry::PositionToIndex, grid_map_geometry.cpp:113-122 */
inline void PositionToIndex(const Grid& g, double px, double py,
                            int& col, int& row)
{
    col = static_cast<int>(std::floor((px - g.offX) / g.res));
    row = static_cast<int>(std::floor((py - g.offY) / g.res));
}

struct ScoreSummary { double normalized, sum, knownRate; };

/* ScorePixelAccurate::Score, score_function_pixel_accurate.cpp:16-58 */
ScoreSummary ScorePixelAccurate(const Grid& g, const Scan& scan,
                                const Pose& sensor)
{
    double sum = 0.0;
    size_t known = 0;
    const size_t n = scan.N();
    for (size_t i = 0; i < n; ++i) {
        double hx, hy;
        HitPoint(scan, sensor, i, hx, hy);
        int col, row;
        PositionToIndex(g, hx, hy, col, row);
        const double p = ProbabilityOr(g, row, col, 0.0);
        if (p == 0.0)
            continue;
        sum += p;
        ++known;
    }
    return ScoreSummary { sum / static_cast<double>(n), sum,
                          static_cast<double>(known) / static_cast<double>(n) };
}

/* ---- sliding-window maximum ----------------------------------------------
 * util.hpp:369-424. out[i] is the maximum of in[s .. s+win-1] with
 * s = min(i, n-win): a forward window whose start is clamped at the far end,
 * so the last `win` outputs are equal. Elements past the end
 * read as 0 (the reference reads them through ValueOr when win > n). */
template <typename In, typename Out>
void SlidingWindowMax(In in, Out out, int n, int win)
{
    /* Monotonic index queue, values decreasing from head to tail */
    std::vector<int> queue(static_cast<size_t>(std::max(n, win)) + 1);
    int head = 0, tail = 0;
    int idxIn = 0, idxOut = 0;
    auto push = [&](int idx) {
        while (tail > head && in(idx) >= in(queue[tail - 1]))
            --tail;
        queue[tail++] = idx;
    };
    for (idxIn = 0; idxIn < win; ++idxIn)
        push(idxIn);
    for (; idxIn < n; ++idxIn) {
        out(idxOut++, in(queue[head]));
        while (tail > head && queue[head] <= idxIn - win)
            ++head;
        push(idxIn);
    }
    for (; idxOut < n; ++idxOut)
        out(idxOut, in(queue[head]));
}

/* PrecomputeGridMap, grid_map_builder.cpp:1015-1065: first the maximum along
 * rows for every column (SlidingWindowMaxRow, :918-949), then along columns
 * for every row (SlidingWindowMaxCol, :952-984). The result is dense (every
 * block allocated, grid_map.cpp:522-535). */
Grid Precompute(const Grid& g, int win)
{
    Grid mid = g;
    mid.allBlocks = true;
    std::fill(mid.cells.begin(), mid.cells.end(), 0);
    for (int c = 0; c < g.cols; ++c)
        SlidingWindowMax(
            [&](int r) { return g.ValueOr(r, c, 0); },
            [&](int r, uint16_t v) { mid.cells[static_cast<size_t>(r) * g.cols + c] = v; },
            g.rows, win);
    Grid out = mid;
    for (int r = 0; r < g.rows; ++r)
        SlidingWindowMax(
            [&](int c) { return mid.ValueOr(r, c, 0); },
            [&](int c, uint16_t v) { out.cells[static_cast<size_t>(r) * g.cols + c] = v; },
            g.cols, win);
    return out;
}

/* PrecomputeGridMaps, grid_map_builder.cpp:987-1012: win = 1, 2, 4 ... */
std::vector<Grid> PrecomputePyramid(const Grid& g, int hmax)
{
    std::vector<Grid> levels;
    levels.reserve(hmax + 1);
    for (int h = 0, win = 1; h <= hmax; ++h, win <<= 1)
        levels.push_back(Precompute(g, win));
    return levels;
}

/* ---- search step / window -------------------------------------------------
 * scan_matcher_correlative.cpp:255-274, scan_matcher_branch_bound.cpp:293-312 */
void SearchStep(const Grid& g, const Scan& scan,
                double& sx, double& sy, double& st)
{
    const double maxRange =
        *std::max_element(scan.ranges.begin(), scan.ranges.end());
    const double theta = g.res / maxRange;
    sx = g.res;
    sy = g.res;
    st = std::acos(1.0 - 0.5 * theta * theta);
}

/* ---- cost function ---------------------------------------------------------
 * cost_function_square_error.cpp: bilinear smoothing (:27-36, :323-347),
 * cost (:48-75), Hessian (:151-195), covariance (:131-146). */
struct MapValues { double dx, dy, m00, m01, m10, m11; };

MapValues ClosestMapValues(const Grid& g, double fx, double fy)
{
    const double x0 = std::floor(fx);
    const double y0 = std::floor(fy);
    const double dx = fx - x0;
    const double dy = fy - y0;
    const int xc0 = std::max(static_cast<int>(x0), 0);
    const int yc0 = std::max(static_cast<int>(y0), 0);
    const int xc1 = std::min(xc0 + 1, g.cols - 1);
    const int yc1 = std::min(yc0 + 1, g.rows - 1);
    return MapValues { dx, dy,
                       ProbabilityOr(g, yc0, xc0, 0.5),
                       ProbabilityOr(g, yc1, xc0, 0.5),
                       ProbabilityOr(g, yc0, xc1, 0.5),
                       ProbabilityOr(g, yc1, xc1, 0.5) };
}

inline double Bilinear(const MapValues& m)
{
    return m.dy * (m.dx * m.m11 + (1.0 - m.dx) * m.m01) +
           (1.0 - m.dy) * (m.dx * m.m10 + (1.0 - m.dx) * m.m00);
}

double CostSquareError(const Grid& g, const Scan& scan, const Pose& sensor)
{
    double cost = 0.0;
    for (size_t i = 0; i < scan.N(); ++i) {
        double hx, hy;
        HitPoint(scan, sensor, i, hx, hy);
        /* PositionToIndexF, grid_map_geometry.cpp:125-133 */
        const double fx = (hx - g.offX) / g.res;
        const double fy = (hy - g.offY) / g.res;
        const double smoothed = Bilinear(ClosestMapValues(g, fx, fy));
        cost += std::pow(1.0 - smoothed, 2.0);
    }
    return cost;
}

void CovarianceSquareError(const Grid& g, const Scan& scan, const Pose& sensor,
                           double scale, double cov[9])
{
    double h[9] = { 0.0 };
    const double invRes = 1.0 / g.res;
    for (size_t i = 0; i < scan.N(); ++i) {
        double hx, hy;
        HitPoint(scan, sensor, i, hx, hy);
        const double fx = (hx - g.offX) / g.res;
        const double fy = (hy - g.offY) / g.res;
        const MapValues m = ClosestMapValues(g, fx, fy);
        const double rx = hx - sensor.x;
        const double ry = hy - sensor.y;
        /* cost_function_square_error.cpp:233-274 */
        const double gx = m.dy * (m.m11 - m.m01) + (1.0 - m.dy) * (m.m10 - m.m00);
        const double gy = m.dx * (m.m11 - m.m10) + (1.0 - m.dx) * (m.m01 - m.m00);
        const double gt = -ry * gx + rx * gy;
        const double grad[3] = { gx * invRes, gy * invRes, gt * invRes };
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                h[r * 3 + c] += grad[r] * grad[c];
    }
    /* inverse (adjugate; the checker compares covariances at 1e-9 relative) */
    const double det =
        h[0] * (h[4] * h[8] - h[5] * h[7]) -
        h[1] * (h[3] * h[8] - h[5] * h[6]) +
        h[2] * (h[3] * h[7] - h[4] * h[6]);
    const double inv = 1.0 / det;
    double a[9];
    a[0] = (h[4] * h[8] - h[5] * h[7]) * inv;
    a[1] = (h[2] * h[7] - h[1] * h[8]) * inv;
    a[2] = (h[1] * h[5] - h[2] * h[4]) * inv;
    a[3] = (h[5] * h[6] - h[3] * h[8]) * inv;
    a[4] = (h[0] * h[8] - h[2] * h[6]) * inv;
    a[5] = (h[2] * h[3] - h[0] * h[5]) * inv;
    a[6] = (h[3] * h[7] - h[4] * h[6]) * inv;
    a[7] = (h[1] * h[6] - h[0] * h[7]) * inv;
    a[8] = (h[0] * h[4] - h[1] * h[3]) * inv;
    for (int i = 0; i < 9; ++i)
        cov[i] = a[i] * scale;
}

/* ---- linear-solver refiner -----------------------------------------------------
 * scan_matcher_linear_solver.cpp:66-170 with cost_function_square_error.cpp:151-195
 * (Hessian and residual). The 3x3 system is solved by a Householder QR with column
 * pivoting (what Eigen's colPivHouseholderQr() does, :161). */
void HessianAndResidual(const Grid& g, const Scan& scan, const Pose& sensor, double h[9], double res[3])
{
    for (int i = 0; i < 9; ++i) h[i] = 0.0;
    for (int i = 0; i < 3; ++i) res[i] = 0.0;
    const double invRes = 1.0 / g.res;
    for (size_t i = 0; i < scan.N(); ++i) {
        double hx, hy;
        HitPoint(scan, sensor, i, hx, hy);
        const double fx = (hx - g.offX) / g.res;
        const double fy = (hy - g.offY) / g.res;
        const MapValues m = ClosestMapValues(g, fx, fy);
        const double rx = hx - sensor.x;
        const double ry = hy - sensor.y;
        const double gx = m.dy * (m.m11 - m.m01) + (1.0 - m.dy) * (m.m10 - m.m00);
        const double gy = m.dx * (m.m11 - m.m10) + (1.0 - m.dx) * (m.m01 - m.m00);
        const double gt = -ry * gx + rx * gy;
        const double grad[3] = { gx * invRes, gy * invRes, gt * invRes };
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                h[r * 3 + c] += grad[r] * grad[c];
        const double mapResidual = 1.0 - Bilinear(m);
        for (int r = 0; r < 3; ++r)
            res[r] += grad[r] * mapResidual;
    }
}

void SolveColPivQr3(const double aIn[9], const double bIn[3], double x[3])
{
    double a[3][3], b[3] = { bIn[0], bIn[1], bIn[2] };
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            a[r][c] = aIn[r * 3 + c];
    int perm[3] = { 0, 1, 2 };
    double maxPivot = 0.0;
    for (int k = 0; k < 3; ++k) {
        int best = k;
        double bestNorm = -1.0;
        for (int c = k; c < 3; ++c) {
            double n2 = 0.0;
            for (int r = k; r < 3; ++r) n2 += a[r][c] * a[r][c];
            if (n2 > bestNorm) { bestNorm = n2; best = c; }
        }
        if (best != k) {
            for (int r = 0; r < 3; ++r) std::swap(a[r][k], a[r][best]);
            std::swap(perm[k], perm[best]);
        }
        double tail = 0.0;
        for (int r = k + 1; r < 3; ++r) tail += a[r][k] * a[r][k];
        const double c0 = a[k][k];
        double beta = c0, tau = 0.0, v[3] = { 0.0, 0.0, 0.0 };
        if (tail > 0.0) {
            beta = std::sqrt(c0 * c0 + tail);
            if (c0 >= 0.0) beta = -beta;
            for (int r = k + 1; r < 3; ++r) v[r] = a[r][k] / (c0 - beta);
            v[k] = 1.0;
            tau = (beta - c0) / beta;
        }
        a[k][k] = beta;
        for (int r = k + 1; r < 3; ++r) a[r][k] = 0.0;
        if (tau != 0.0) {
            for (int c = k + 1; c < 3; ++c) {
                double d = 0.0;
                for (int r = k; r < 3; ++r) d += v[r] * a[r][c];
                for (int r = k; r < 3; ++r) a[r][c] -= tau * v[r] * d;
            }
            double d = 0.0;
            for (int r = k; r < 3; ++r) d += v[r] * b[r];
            for (int r = k; r < 3; ++r) b[r] -= tau * v[r] * d;
        }
        maxPivot = std::max(maxPivot, std::fabs(beta));
    }
    const double threshold = 2.220446049250313e-16 * 3.0 * maxPivot;
    int rank = 0;
    for (int k = 0; k < 3; ++k)
        if (std::fabs(a[k][k]) > threshold) ++rank;
    double y[3] = { 0.0, 0.0, 0.0 };
    for (int k = rank - 1; k >= 0; --k) {
        double acc = b[k];
        for (int c = k + 1; c < rank; ++c) acc -= a[k][c] * y[c];
        y[k] = acc / a[k][k];
    }
    for (int k = 0; k < 3; ++k) x[perm[k]] = y[k];
}

constexpr double kCovarianceScale = 1e4;

/* ScanMatcherLinearSolver::OptimizePose; `lambda` is the matcher's member mLambda */
int RefinePose(const Grid& g, const Scan& scan, const Pose& init, int iterationsMax,
               double convergenceThreshold, double& lambda, Pose& estimated, double& cost,
               double cov[9])
{
    const Pose sensor = Compound(init, scan.rel);
    const double initialCost = CostSquareError(g, scan, sensor);
    double prevCost = initialCost;
    Pose best = sensor;
    int iterations = 0;
    while (true) {
        double h[9], res[3], d[3];
        HessianAndResidual(g, scan, best, h, res);
        h[0] += lambda; h[4] += lambda; h[8] += lambda;
        SolveColPivQr3(h, res, d);
        best = Pose { best.x + d[0], best.y + d[1], best.t + d[2] };
        cost = CostSquareError(g, scan, best);
        if (++iterations >= iterationsMax || std::fabs(prevCost - cost) < convergenceThreshold)
            break;
        if (cost < prevCost)
            lambda = std::max(1e-8, lambda * 0.5);
        else
            lambda = std::min(1e-4, lambda * 2.0);
        prevCost = cost;
    }
    estimated = MoveBackward(best, scan.rel);
    CovarianceSquareError(g, scan, best, kCovarianceScale, cov);
    return iterations;
}

/* Epilogue shared by all matchers (scan_matcher_correlative.cpp:203-219) */
void Epilogue(const Grid& g, const Scan& scan, const Pose& best, orc_result* out)
{
    out->best_sensor_pose[0] = best.x;
    out->best_sensor_pose[1] = best.y;
    out->best_sensor_pose[2] = best.t;
    out->norm_cost = CostSquareError(g, scan, best) /
                     static_cast<double>(scan.N());
    const Pose est = MoveBackward(best, scan.rel);
    out->est_pose[0] = est.x;
    out->est_pose[1] = est.y;
    out->est_pose[2] = est.t;
    CovarianceSquareError(g, scan, best, kCovarianceScale, out->cov);
}

void IntegerScoreAt(const Grid& g, const Scan& scan, const Pose& pose,
                    orc_result* out)
{
    int64_t sum = 0;
    int known = 0;
    for (size_t i = 0; i < scan.N(); ++i) {
        double hx, hy;
        HitPoint(scan, pose, i, hx, hy);
        int col, row;
        PositionToIndex(g, hx, hy, col, row);
        const uint16_t v = g.ValueOr(row, col, 0);
        if (v != 0) { sum += v; ++known; }
    }
    out->sum_value = sum;
    out->n_known = known;
}

/* ---- real-time correlative matcher ----------------------------------------
 * scan_matcher_correlative.cpp:118-244, :277-368 */
struct Idx { int col, row; };

ScoreSummary ScoreShifted(const Grid& g, const std::vector<Idx>& idx,
                          int offX, int offY, int64_t* sumValue, int* nKnown)
{
    size_t known = 0;
    double sum = 0.0;
    int64_t sv = 0;
    for (const Idx& p : idx) {
        const double prob = ProbabilityOr(g, p.row + offY, p.col + offX, 0.0);
        if (prob == 0.0)
            continue;
        sum += prob;
        sv += g.ValueOr(p.row + offY, p.col + offX, 0);
        ++known;
    }
    if (sumValue) *sumValue = sv;
    if (nKnown) *nKnown = static_cast<int>(known);
    return ScoreSummary { sum / static_cast<double>(idx.size()), sum,
                          static_cast<double>(known) /
                          static_cast<double>(idx.size()) };
}

void MatchRT(const Grid& g, const Grid& coarse, const Scan& scan,
             const Pose& init, int lowRes, double rangeX, double rangeY,
             double rangeT, double scoreThr, double knownThr, orc_result* out)
{
    const Pose sensor = Compound(init, scan.rel);
    double stepX, stepY, stepT;
    SearchStep(g, scan, stepX, stepY, stepT);
    const int winX = static_cast<int>(std::ceil(0.5 * rangeX / stepX));
    const int winY = static_cast<int>(std::ceil(0.5 * rangeY / stepY));
    const int winT = static_cast<int>(std::ceil(0.5 * rangeT / stepT));

    double scoreMax = scoreThr;
    int bestX = -winX, bestY = -winY, bestT = -winT;
    int ignored = 0, processed = 0;
    std::vector<Idx> idx(scan.N());

    for (int t = -winT; t <= winT; ++t) {
        const Pose pose { sensor.x, sensor.y, sensor.t + stepT * t };
        for (size_t i = 0; i < scan.N(); ++i) {
            double hx, hy;
            HitPoint(scan, pose, i, hx, hy);
            PositionToIndex(coarse, hx, hy, idx[i].col, idx[i].row);
        }
        for (int x = -winX; x <= winX; x += lowRes) {
            for (int y = -winY; y <= winY; y += lowRes) {
                const ScoreSummary c =
                    ScoreShifted(coarse, idx, x, y, nullptr, nullptr);
                if (c.normalized <= scoreMax || c.knownRate <= knownThr) {
                    ++ignored;
                    continue;
                }
                for (int fx = x; fx < x + lowRes; ++fx)
                    for (int fy = y; fy < y + lowRes; ++fy) {
                        const ScoreSummary f =
                            ScoreShifted(g, idx, fx, fy, nullptr, nullptr);
                        if (scoreMax < f.normalized) {
                            scoreMax = f.normalized;
                            bestX = fx; bestY = fy; bestT = t;
                        }
                    }
                ++processed;
            }
        }
    }

    *out = orc_result { };
    out->found = scoreMax > scoreThr ? 1 : 0;
    out->best_x = bestX; out->best_y = bestY; out->best_t = bestT;
    out->win_x = winX; out->win_y = winY; out->win_t = winT;
    out->step_x = stepX; out->step_y = stepY; out->step_t = stepT;
    out->n_processed = processed;
    out->n_ignored = ignored;

    /* Diagnostics at the best pose (indices of that angle, integer shift) */
    const Pose anglePose { sensor.x, sensor.y, sensor.t + stepT * bestT };
    for (size_t i = 0; i < scan.N(); ++i) {
        double hx, hy;
        HitPoint(scan, anglePose, i, hx, hy);
        PositionToIndex(g, hx, hy, idx[i].col, idx[i].row);
    }
    const ScoreSummary s =
        ScoreShifted(g, idx, bestX, bestY, &out->sum_value, &out->n_known);
    out->score = s.normalized;
    out->known_rate = s.knownRate;

    const Pose best { sensor.x + bestX * stepX, sensor.y + bestY * stepY,
                      sensor.t + bestT * stepT };
    Epilogue(g, scan, best, out);
}

/* ---- branch-and-bound matcher -----------------------------------------------
 * scan_matcher_branch_bound.cpp:111-278; Node: scan_matcher_branch_bound.hpp:67-106
 * (ordering by normalized score only; std::priority_queue as in the reference
 * so that equal-score ties break the same way under the same libstdc++). */
struct Node
{
    int x, y, t, h;
    double score, knownRate;
    bool operator<(const Node& o) const { return score < o.score; }
};

void MatchBB(const Grid& g, const std::vector<Grid>& pyramid, const Scan& scan,
             const Pose& init, int hmax, double rangeX, double rangeY,
             double rangeT, double scoreThr, double knownThr, orc_result* out)
{
    const Pose sensor = Compound(init, scan.rel);
    double stepX, stepY, stepT;
    SearchStep(g, scan, stepX, stepY, stepT);
    const int winX = static_cast<int>(std::ceil(0.5 * rangeX / stepX));
    const int winY = static_cast<int>(std::ceil(0.5 * rangeY / stepY));
    const int winT = static_cast<int>(std::ceil(0.5 * rangeT / stepT));

    double scoreMax = scoreThr;
    int bestX = 0, bestY = 0, bestT = 0;
    int ignored = 0, processed = 0;
    std::priority_queue<Node> queue;
    const int winMax = 1 << hmax;

    auto append = [&](int x, int y, int t, int h) {
        const Pose pose { sensor.x + x * stepX, sensor.y + y * stepY,
                          sensor.t + t * stepT };
        const ScoreSummary s = ScorePixelAccurate(pyramid.at(h), scan, pose);
        if (s.normalized > scoreMax)
            queue.push(Node { x, y, t, h, s.normalized, s.knownRate });
        if (s.normalized <= scoreMax)
            ++ignored;
    };

    for (int x = -winX; x <= winX; x += winMax)
        for (int y = -winY; y <= winY; y += winMax)
            for (int t = -winT; t <= winT; ++t)
                append(x, y, t, hmax);

    while (!queue.empty()) {
        const Node cur = queue.top();
        if (cur.score <= scoreMax || cur.knownRate <= knownThr) {
            queue.pop();
            ++ignored;
            continue;
        }
        queue.pop();
        ++processed;
        if (cur.h == 0) {
            bestX = cur.x; bestY = cur.y; bestT = cur.t;
            scoreMax = cur.score;
        } else {
            const int h = cur.h - 1;
            const int w = 1 << h;
            append(cur.x, cur.y, cur.t, h);
            append(cur.x + w, cur.y, cur.t, h);
            append(cur.x, cur.y + w, cur.t, h);
            append(cur.x + w, cur.y + w, cur.t, h);
        }
    }

    *out = orc_result { };
    out->found = scoreMax > scoreThr ? 1 : 0;
    out->best_x = bestX; out->best_y = bestY; out->best_t = bestT;
    out->win_x = winX; out->win_y = winY; out->win_t = winT;
    out->step_x = stepX; out->step_y = stepY; out->step_t = stepT;
    out->n_processed = processed;
    out->n_ignored = ignored;

    const Pose best { sensor.x + stepX * bestX, sensor.y + stepY * bestY,
                      sensor.t + stepT * bestT };
    const ScoreSummary s = ScorePixelAccurate(g, scan, best);
    out->score = s.normalized;
    out->known_rate = s.knownRate;
    IntegerScoreAt(g, scan, best, out);
    Epilogue(g, scan, best, out);
}

/* ---- exhaustive grid search ---------------------------------------------------
 * scan_matcher_grid_search.cpp:84-178 (loops with accumulating doubles,
 * order dy -> dx -> dtheta, :118-120) */
void MatchGrid(const Grid& g, const Scan& scan, const Pose& init,
               double rangeX, double rangeY, double rangeT,
               double stepX, double stepY, double stepT,
               double scoreThr, double knownThr, orc_result* out)
{
    const Pose sensor = Compound(init, scan.rel);
    const double rx = rangeX / 2.0, ry = rangeY / 2.0, rt = rangeT / 2.0;
    double scoreMax = scoreThr;
    int evaluations = 0, updates = 0;
    Pose best = sensor;
    int bestIx = -1, bestIy = -1, bestIt = -1;

    int iy = 0;
    for (double dy = -ry; dy <= ry; dy += stepY, ++iy) {
        int ix = 0;
        for (double dx = -rx; dx <= rx; dx += stepX, ++ix) {
            int it = 0;
            for (double dt = -rt; dt <= rt; dt += stepT, ++it) {
                const Pose pose { sensor.x + dx, sensor.y + dy, sensor.t + dt };
                const ScoreSummary s = ScorePixelAccurate(g, scan, pose);
                ++evaluations;
                if (s.normalized > scoreMax && s.knownRate > knownThr) {
                    scoreMax = s.normalized;
                    best = pose;
                    bestIx = ix; bestIy = iy; bestIt = it;
                    ++updates;
                }
            }
        }
    }

    *out = orc_result { };
    out->found = scoreMax > scoreThr ? 1 : 0;
    out->best_x = bestIx; out->best_y = bestIy; out->best_t = bestIt;
    out->step_x = stepX; out->step_y = stepY; out->step_t = stepT;
    out->n_processed = evaluations;
    out->n_ignored = updates;
    const ScoreSummary s = ScorePixelAccurate(g, scan, best);
    out->score = s.normalized;
    out->known_rate = s.knownRate;
    IntegerScoreAt(g, scan, best, out);
    Epilogue(g, scan, best, out);
}

Scan MakeScan(const double* angles, const double* ranges, int n,
              const double rel[3])
{
    Scan s;
    s.angles.assign(angles, angles + n);
    s.ranges.assign(ranges, ranges + n);
    s.rel = Pose { rel[0], rel[1], rel[2] };
    return s;
}

/* ---- loop detector ---------------------------------------------------------------
 * loop_detector_branch_bound.cpp:59-156 with a pass-through final matcher */
struct LoopDetector
{
    int hmax;
    double rangeX, rangeY, rangeT, scoreThr, knownThr;
    int nThreads;
    /* final matcher: pass-through (default) or the linear-solver refiner; its damping
     * factor is per detector (thread) and carries over from query to query */
    bool linearSolver = false;
    int finalIterations = 10;
    double finalConvergence = 1e-4;
    std::vector<double> finalLambda;
    /* one pyramid cache per thread, keyed by local map id (never evicted,
     * loop_detector_branch_bound.cpp:83-89) */
    std::vector<std::map<int, std::vector<Grid>>> caches;
};

} /* namespace */

extern "C" {

const char* orc_kind(void) { return "port"; }

void* orc_grid_create(const uint16_t* dense, int rows, int cols,
                      double resolution, double offset_x, double offset_y)
{
    if (rows <= 0 || cols <= 0 || rows % kBlock || cols % kBlock)
        return nullptr;
    auto* g = new Grid;
    g->rows = rows; g->cols = cols;
    g->res = resolution; g->offX = offset_x; g->offY = offset_y;
    g->cells.assign(dense, dense + static_cast<size_t>(rows) * cols);
    const int bc = cols >> kLog2Block;
    g->blockAllocated.assign(static_cast<size_t>(rows >> kLog2Block) * bc, 0);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c)
            if (g->cells[static_cast<size_t>(r) * cols + c] != 0)
                g->blockAllocated[(r >> kLog2Block) * bc + (c >> kLog2Block)] = 1;
    return g;
}

void orc_grid_destroy(void* grid) { delete static_cast<Grid*>(grid); }

int orc_precompute(void* grid, int win, uint16_t* out)
{
    const Grid p = Precompute(*static_cast<Grid*>(grid), win);
    std::copy(p.cells.begin(), p.cells.end(), out);
    return 0;
}

int orc_precompute_pyramid(void* grid, int hmax, uint16_t* out)
{
    const Grid& g = *static_cast<Grid*>(grid);
    const std::vector<Grid> levels = PrecomputePyramid(g, hmax);
    const size_t cells = static_cast<size_t>(g.rows) * g.cols;
    for (size_t h = 0; h < levels.size(); ++h)
        std::copy(levels[h].cells.begin(), levels[h].cells.end(),
                  out + h * cells);
    return 0;
}

int orc_match_rt(void* grid, const double* angles, const double* ranges, int n,
                 const double init_pose[3], const double rel_sensor_pose[3],
                 int low_res, double range_x, double range_y, double range_t,
                 double score_thr, double known_thr, orc_result* out)
{
    const Grid& g = *static_cast<Grid*>(grid);
    const Scan scan = MakeScan(angles, ranges, n, rel_sensor_pose);
    const Grid coarse = Precompute(g, low_res);
    MatchRT(g, coarse, scan, Pose { init_pose[0], init_pose[1], init_pose[2] },
            low_res, range_x, range_y, range_t, score_thr, known_thr, out);
    return 0;
}

int orc_match_bb(void* grid, const double* angles, const double* ranges, int n,
                 const double init_pose[3], const double rel_sensor_pose[3],
                 int hmax, double range_x, double range_y, double range_t,
                 double score_thr, double known_thr, orc_result* out)
{
    const Grid& g = *static_cast<Grid*>(grid);
    const Scan scan = MakeScan(angles, ranges, n, rel_sensor_pose);
    const std::vector<Grid> pyramid = PrecomputePyramid(g, hmax);
    MatchBB(g, pyramid, scan, Pose { init_pose[0], init_pose[1], init_pose[2] },
            hmax, range_x, range_y, range_t, score_thr, known_thr, out);
    return 0;
}

int orc_match_grid(void* grid, const double* angles, const double* ranges, int n,
                   const double init_pose[3], const double rel_sensor_pose[3],
                   double range_x, double range_y, double range_t,
                   double step_x, double step_y, double step_t,
                   double score_thr, double known_thr, orc_result* out)
{
    const Grid& g = *static_cast<Grid*>(grid);
    const Scan scan = MakeScan(angles, ranges, n, rel_sensor_pose);
    MatchGrid(g, scan, Pose { init_pose[0], init_pose[1], init_pose[2] },
              range_x, range_y, range_t, step_x, step_y, step_t,
              score_thr, known_thr, out);
    return 0;
}

void* orc_loopdet_create(int hmax, double range_x, double range_y, double range_t,
                         double score_thr, double known_thr, int n_threads)
{
    auto* det = new LoopDetector;
    det->hmax = hmax;
    det->rangeX = range_x; det->rangeY = range_y; det->rangeT = range_t;
    det->scoreThr = score_thr; det->knownThr = known_thr;
    det->nThreads = std::max(1, n_threads);
    det->caches.resize(det->nThreads);
    det->finalLambda.assign(det->nThreads, 1e-4);
    return det;
}

void orc_loopdet_use_linear_solver(void* detPtr, int iterations_max, double convergence_threshold,
                                   double initial_lambda)
{
    auto* det = static_cast<LoopDetector*>(detPtr);
    det->linearSolver = true;
    det->finalIterations = iterations_max;
    det->finalConvergence = convergence_threshold;
    det->finalLambda.assign(det->nThreads, initial_lambda);
}

int orc_refine(void* grid, const double* angles, const double* ranges, int n,
               const double init_pose[3], const double rel_sensor_pose[3],
               int iterations_max, double convergence_threshold, double* lambda,
               orc_result* out)
{
    const Grid& g = *static_cast<Grid*>(grid);
    const Scan scan = MakeScan(angles, ranges, n, rel_sensor_pose);
    const Pose init { init_pose[0], init_pose[1], init_pose[2] };
    const orc_result empty { };
    *out = empty;
    Pose est;
    double cost = 0.0;
    out->n_processed = RefinePose(g, scan, init, iterations_max, convergence_threshold, *lambda,
                                  est, cost, out->cov);
    out->found = 1;
    out->est_pose[0] = est.x; out->est_pose[1] = est.y; out->est_pose[2] = est.t;
    out->norm_cost = cost / static_cast<double>(scan.N());
    return 0;
}

void orc_loopdet_destroy(void* det) { delete static_cast<LoopDetector*>(det); }

void orc_loopdet_clear_cache(void* detPtr)
{
    auto* det = static_cast<LoopDetector*>(detPtr);
    for (auto& c : det->caches)
        c.clear();
}

int orc_loopdet_detect(void* detPtr, int n_queries,
                       void* const* grids, const int32_t* map_ids,
                       const double* map_global_poses,
                       const int32_t* scan_idx, const double* scan_global_poses,
                       int n_scans, int n_beams,
                       const double* angles, const double* ranges,
                       orc_result* out, double* elapsed_s)
{
    auto* det = static_cast<LoopDetector*>(detPtr);
    const double rel[3] = { 0.0, 0.0, 0.0 };
    std::vector<Scan> scans;
    for (int s = 0; s < n_scans; ++s)
        scans.push_back(MakeScan(angles + static_cast<size_t>(s) * n_beams,
                                 ranges + static_cast<size_t>(s) * n_beams,
                                 n_beams, rel));
    std::vector<double> times(det->nThreads, 0.0);

    auto worker = [&](int t) {
        const int begin = static_cast<int>(
            static_cast<long long>(n_queries) * t / det->nThreads);
        const int end = static_cast<int>(
            static_cast<long long>(n_queries) * (t + 1) / det->nThreads);
        const auto t0 = std::chrono::steady_clock::now();
        for (int q = begin; q < end; ++q) {
            const Grid& g = *static_cast<Grid*>(grids[q]);
            auto& cache = det->caches[t];
            auto it = cache.find(map_ids[q]);
            if (it == cache.end())
                it = cache.emplace(map_ids[q],
                                   PrecomputePyramid(g, det->hmax)).first;
            const Pose mapPose { map_global_poses[3 * q],
                map_global_poses[3 * q + 1], map_global_poses[3 * q + 2] };
            const Pose scanPose { scan_global_poses[3 * q],
                scan_global_poses[3 * q + 1], scan_global_poses[3 * q + 2] };
            const Pose init = InverseCompound(mapPose, scanPose);
            MatchBB(g, it->second, scans[scan_idx[q]], init, det->hmax,
                    det->rangeX, det->rangeY, det->rangeT,
                    det->scoreThr, det->knownThr, &out[q]);
            if (!out[q].found) {
                /* the reference emits nothing for this query */
                const orc_result empty { };
                out[q] = empty;
            } else if (det->linearSolver) {
                /* loop_detector_branch_bound.cpp:110-135: the refined pose and its covariance
                 * replace the coarse ones */
                const Pose coarse { out[q].est_pose[0], out[q].est_pose[1], out[q].est_pose[2] };
                Pose est;
                double cost = 0.0;
                RefinePose(g, scans[scan_idx[q]], coarse, det->finalIterations, det->finalConvergence,
                           det->finalLambda[t], est, cost, out[q].cov);
                out[q].est_pose[0] = est.x; out[q].est_pose[1] = est.y; out[q].est_pose[2] = est.t;
            }
        }
        const auto t1 = std::chrono::steady_clock::now();
        times[t] = std::chrono::duration<double>(t1 - t0).count();
    };

    if (det->nThreads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> threads;
        for (int t = 0; t < det->nThreads; ++t)
            threads.emplace_back(worker, t);
        for (auto& th : threads)
            th.join();
    }
    if (elapsed_s != nullptr)
        *elapsed_s = *std::max_element(times.begin(), times.end());
    return 0;
}

} /* extern "C" */
