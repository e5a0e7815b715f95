/* pose_graph.hpp -- the pose graph the SLAM loop keeps on the host, in the shape of the reference's
 * (mapping/pose_graph.hpp, pose_graph_node.hpp, pose_graph_edge.hpp): scan nodes, local-map nodes and the
 * edges between them. Ids are consecutive from 0 like the reference assigns them
 * (grid_map_builder.cpp:216-218, 303-304), so an id is also the position in its vector.
 * Nothing here touches the device. */
#pragma once

#include <array>
#include <cmath>
#include <vector>

#include "csm_host/types.hpp"

namespace csm_host {

using Mat3 = std::array<double, 9>;         /* row-major 3 x 3 */

/* pose_graph_node.hpp:41-73 */
struct ScanNode
{
    int node_id = 0;
    int local_map_id = 0;
    Pose2D local_pose;                      /* in the frame of its local map */
    ScanDataPtr scan;
    Pose2D global_pose;
};

/* pose_graph_node.hpp:20-38 */
struct LocalMapNode
{
    int local_map_id = 0;
    Pose2D global_pose;
};

enum class EdgeType { IntraLocalMap, InterLocalMap };          /* pose_graph_edge.hpp:19-23 */
enum class ConstraintType { Odometry, Loop };                  /* pose_graph_edge.hpp:26-30 */

/* pose_graph_edge.hpp:33-77: always local map node -> scan node */
struct PoseGraphEdge
{
    int local_map_id = 0;
    int scan_node_id = 0;
    EdgeType edge_type = EdgeType::IntraLocalMap;
    ConstraintType constraint_type = ConstraintType::Odometry;
    Pose2D relative_pose;
    Mat3 information {};
    bool IsOdometryConstraint() const { return constraint_type == ConstraintType::Odometry; }
    bool IsLoopClosingConstraint() const { return constraint_type == ConstraintType::Loop; }
};

struct PoseGraph
{
    std::vector<ScanNode> scan_nodes;
    std::vector<LocalMapNode> local_map_nodes;
    std::vector<PoseGraphEdge> edges;
};

/* What the optimiser sees of an edge (pose_graph.hpp, EdgePose): indices into the two pose vectors */
struct EdgePose
{
    bool is_loop_closing = false;
    int local_map_index = 0;
    int scan_node_index = 0;
    std::array<double, 3> relative_pose {};
    Mat3 information {};
};

/* The seam to the pose-graph optimiser (mapping/pose_graph_optimizer.hpp:14-29): the poses of the finished
 * local maps and their scan nodes go in and come back adjusted; the edges are constants. The reference's
 * implementations (g2o, pose_graph_optimizer_g2o.cpp:50-180, and its own Levenberg-Marquardt) stay on the
 * CPU and plug in here unchanged; this package ships none of them. */
class PoseGraphOptimizer
{
public:
    virtual ~PoseGraphOptimizer() = default;
    virtual void Optimize(std::vector<std::array<double, 3>>& local_map_poses,
                          std::vector<std::array<double, 3>>& scan_poses,
                          const std::vector<EdgePose>& edges) = 0;
};

/* Leaves every pose where it is: the stand-in behind the seam when no optimiser is linked */
class PoseGraphOptimizerIdentity final : public PoseGraphOptimizer
{
public:
    void Optimize(std::vector<std::array<double, 3>>&, std::vector<std::array<double, 3>>&,
                  const std::vector<EdgePose>& edges) override { mLastNumOfEdges = static_cast<int>(edges.size()); ++mCalls; }
    int Calls() const { return mCalls; }
    int LastNumOfEdges() const { return mLastNumOfEdges; }

private:
    int mCalls = 0, mLastNumOfEdges = 0;
};

/* util.hpp:28, 280-301 */
constexpr double kPi = 3.14159265358979323846;
inline double NormalizeAngle(double theta)
{
    double t = std::fmod(theta, 2.0 * kPi);
    if (t > kPi) t -= 2.0 * kPi;
    else if (t < -kPi) t += 2.0 * kPi;
    return t;
}
inline Pose2D NormalizeAngle(const Pose2D& p) { return Pose2D { p.x, p.y, NormalizeAngle(p.theta) }; }
/* pose.hpp:124-136 */
inline double Distance(const Pose2D& p) { return std::hypot(p.x, p.y); }
inline double Distance(const Pose2D& a, const Pose2D& b) { return std::hypot(a.x - b.x, a.y - b.y); }

Mat3 Multiply(const Mat3& a, const Mat3& b);
Mat3 Transpose(const Mat3& a);
Mat3 Inverse(const Mat3& a);
/* util.hpp:320-352: R(angle) * cov * R(angle)^T */
Mat3 RotateCovariance(double angle, const Mat3& cov);
inline Mat3 ConvertCovarianceFromWorldToLocal(const Pose2D& pose, const Mat3& cov) { return RotateCovariance(-pose.theta, cov); }
inline Mat3 ConvertCovarianceFromLocalToWorld(const Pose2D& pose, const Mat3& cov) { return RotateCovariance(pose.theta, cov); }

} /* namespace csm_host */
