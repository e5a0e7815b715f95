/* carmen_log.hpp -- the on-disk formats either side of the full loop.
 *
 *   CarmenLogReader   the reference's input format: io/carmen/carmen_reader.cpp:11-535. A log is text, one record
 *                     per line, the first token names the record: PARAM, ODOM, RAWLASER1..4, ROBOTLASER1..2,
 *                     FLASER / RLASER (old format, with poses), LASER3 / LASER4 (old format, ranges only); every
 *                     other record is skipped. Records come back in file order as CarmenRecord (the reference's
 *                     Sensor::OdometryData / Sensor::ScanData, sensor/sensor_data.hpp:34-178).
 *   CarmenLogWriter   ODOM / ROBOTLASER1 / FLASER records with 17 significant digits, so that a synthetic run
 *                     written and read back is the same run bit for bit (the reference has no writer).
 *   WriteMetricsJson  the reference's `<output>.metric.json` (slam_launcher.cpp:171-181, metric/metric.cpp:460-496,
 *                     metric.hpp:32-59, 611-621): the value sequences a MetricRecorder holds, numbers as strings with
 *                     six decimals, the four other metric families present and empty.
 */
#pragma once

#include <iosfwd>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "csm_host/metrics.hpp"
#include "csm_host/types.hpp"

namespace csm_host {

struct CarmenRecord
{
    enum class Kind { Odometry = 0, Scan = 1 };
    Kind kind = Kind::Odometry;
    std::string sensor_id;
    double time_stamp = 0.0;              /* SensorData::TimeStamp: the record's IPC time stamp */
    Pose2D odom_pose;                     /* OdometryData::Pose / ScanData::OdomPose */
    Pose2D velocity;                      /* (forward, 0, angular) */
    std::shared_ptr<ScanData> scan;       /* Kind::Scan: angles, ranges, sensor pose on the robot, range limits */
    double min_angle = 0.0, max_angle = 0.0;
};

class CarmenLogReader
{
public:
    /* carmen_reader.cpp:11-43: clears `records`, returns true (like the reference, malformed lines are not an
     * error; a record cut short keeps zeros in the fields it lacks) */
    bool Load(std::istream& input, std::vector<CarmenRecord>& records);
    bool LoadFile(const std::string& path, std::vector<CarmenRecord>& records);
    /* the PARAM records seen by the last Load (first value wins, :113-132) */
    const std::unordered_map<std::string, std::string>& Parameters() const { return mParams; }

    /* carmen_reader.cpp:463-505 */
    static double GuessAngleRange(int num_readings);
    static double GuessAngleIncrement(int num_readings);

private:
    std::unordered_map<std::string, std::string> mParams;
};

class CarmenLogWriter
{
public:
    explicit CarmenLogWriter(std::ostream& out, const std::string& host = "b200") : mOut(out), mHost(host) { }
    void Param(const std::string& name, const std::string& value);
    void Odom(const Pose2D& pose, double forward_velocity, double angular_velocity, double time_stamp);
    /* ROBOTLASER1: evenly spaced beams from start_angle; laser and robot pose in the odometry frame */
    void RobotLaser(const std::string& sensor_id, double start_angle, double angular_resolution, double max_range,
                    const std::vector<double>& ranges, const Pose2D& laser_pose, const Pose2D& robot_pose,
                    double time_stamp);
    /* FLASER: the old format; the beam geometry comes from PARAM records or the reader's guesses */
    void OldLaser(const std::string& sensor_id, const std::vector<double>& ranges, const Pose2D& laser_pose,
                  const Pose2D& robot_pose, double time_stamp);

private:
    void Tail(double time_stamp);
    std::ostream& mOut;
    std::string mHost;
};

/* the "Values" string of one value sequence: the reference's VecToString on the sequence's element type */
std::string MetricValuesToString(const std::string& id, const std::vector<double>& values);
/* slam_launcher.cpp:171-181 */
void WriteMetricsJson(std::ostream& out, const MetricRecorder& metrics);
bool SaveMetrics(const std::string& output_path, const MetricRecorder& metrics);   /* writes output_path + ".metric.json" */

} /* namespace csm_host */
