/* carmen_log.cpp -- Carmen log reader / writer and the metrics file (carmen_log.hpp).
 *
 * The reader follows io/carmen/carmen_reader.cpp of the reference record by record; what a field means and
 * which defaults apply is cited at each parser. Numbers are converted with strtod / strtol on whitespace
 * separated tokens, which is what the reference's `stream >> double` does underneath (num_get -> strtod):
 * the same text gives the same double. */
#include "csm_host/carmen_log.hpp"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <istream>
#include <ostream>

namespace csm_host {

namespace {

constexpr double kPi = 3.14159265358979323846;          /* util.hpp:27-32 */
constexpr double kPiHalf = 1.57079632679489661923;

/* whitespace separated tokens of one line; past the end every field reads as zero / empty */
class Fields
{
public:
    explicit Fields(const std::string& line) : mP(line.c_str()) { }
    bool Next(const char*& begin, const char*& end)
    {
        while (*mP == ' ' || *mP == '\t' || *mP == '\r' || *mP == '\n' || *mP == '\v' || *mP == '\f') ++mP;
        if (!*mP) return false;
        begin = mP;
        while (*mP && !(*mP == ' ' || *mP == '\t' || *mP == '\r' || *mP == '\n' || *mP == '\v' || *mP == '\f')) ++mP;
        end = mP;
        return true;
    }
    std::string Str()
    {
        const char *b, *e;
        return Next(b, e) ? std::string(b, e) : std::string();
    }
    double Real()
    {
        const char *b, *e;
        return Next(b, e) ? std::strtod(b, nullptr) : 0.0;
    }
    int Int()
    {
        const char *b, *e;
        return Next(b, e) ? static_cast<int>(std::strtol(b, nullptr, 10)) : 0;
    }
    Pose2D Pose()
    {
        Pose2D p;
        p.x = Real(); p.y = Real(); p.theta = Real();
        return p;
    }
    void Ranges(int n, std::vector<double>& out)
    {
        out.clear();
        if (n > 0) out.reserve(static_cast<std::size_t>(n));
        for (int i = 0; i < n; ++i) out.push_back(Real());
    }

private:
    const char* mP;
};

enum class RecordType { None, Param, Odom, RawLaser, RobotLaser, OldLaser, OldOtherLaser };

/* carmen_reader.cpp:507-530 (TRUEPOS is named there and then ignored, :105-108) */
RecordType TypeOf(const std::string& id)
{
    if (id == "PARAM") return RecordType::Param;
    if (id == "ODOM") return RecordType::Odom;
    if (id == "RAWLASER1" || id == "RAWLASER2" || id == "RAWLASER3" || id == "RAWLASER4") return RecordType::RawLaser;
    if (id == "ROBOTLASER1" || id == "ROBOTLASER2") return RecordType::RobotLaser;
    if (id == "FLASER" || id == "RLASER") return RecordType::OldLaser;
    if (id == "LASER3" || id == "LASER4") return RecordType::OldOtherLaser;
    return RecordType::None;
}

void EvenAngles(double start, double increment, int n, std::vector<double>& angles)
{
    angles.clear();
    if (n > 0) angles.reserve(static_cast<std::size_t>(n));
    for (int i = 0; i < n; ++i) angles.push_back(start + increment * i);       /* :225-228 */
}

/* beam geometry of the old formats: PARAM records first, the guesses otherwise (:355-377, 420-442) */
struct OldGeometry { double min_range, max_range, increment, min_angle, max_angle; };

OldGeometry OldLaserGeometry(const std::unordered_map<std::string, std::string>& params, int n)
{
    auto find = [&](const char* name, double& value) {
        const auto it = params.find(name);
        if (it == params.end()) return false;
        value = std::stod(it->second);
        return true;
    };
    OldGeometry g { 0.0, 80.0, 0.0, -kPiHalf, 0.0 };
    find("Laser.MinRange", g.min_range);
    find("Laser.MaxRange", g.max_range);
    const bool has_increment = find("Laser.AngleIncrement", g.increment);
    if (!has_increment) g.increment = CarmenLogReader::GuessAngleIncrement(n);
    find("Laser.MinAngle", g.min_angle);
    if (!find("Laser.MaxAngle", g.max_angle))
        g.max_angle = has_increment ? g.min_angle + g.increment * static_cast<double>(n)
                                    : g.min_angle + CarmenLogReader::GuessAngleRange(n);
    return g;
}

} /* namespace */

/* carmen_reader.cpp:463-483 */
double CarmenLogReader::GuessAngleRange(int n)
{
    switch (n) {
        case 180: return kPi * 179.0 / 180.0;
        case 360: return kPi * 179.5 / 180.0;
        case 401: return kPi * 100.0 / 180.0;
        case 400: return kPi * 99.75 / 180.0;
        default:  return kPi;                            /* 181, 361 and everything unknown */
    }
}

/* carmen_reader.cpp:485-505 */
double CarmenLogReader::GuessAngleIncrement(int n)
{
    switch (n) {
        case 180: case 181: return kPi / 180.0;
        case 360: case 361: return kPi / 360.0;
        case 400: case 401: return kPi / 720.0;
        default:            return GuessAngleRange(n) / static_cast<double>(n - 1);
    }
}

bool CarmenLogReader::Load(std::istream& input, std::vector<CarmenRecord>& records)
{
    records.clear();
    mParams.clear();
    std::string line;
    while (std::getline(input, line)) {
        Fields f(line);
        CarmenRecord rec;
        rec.sensor_id = f.Str();
        switch (TypeOf(rec.sensor_id)) {
            case RecordType::Param: {                    /* :113-132, insert: the first value of a name stays */
                const std::string name = f.Str();
                mParams.insert(std::make_pair(name, f.Str()));
                break;
            }
            case RecordType::Odom: {                     /* :136-160: x y theta tv rv accel, then the header */
                rec.kind = CarmenRecord::Kind::Odometry;
                rec.odom_pose = f.Pose();
                rec.velocity.x = f.Real();
                rec.velocity.theta = f.Real();
                f.Real();                                /* acceleration */
                rec.time_stamp = f.Real();
                records.push_back(std::move(rec));
                break;
            }
            case RecordType::RawLaser:                   /* :164-236 */
            case RecordType::RobotLaser: {               /* :240-317 */
                const bool robot = TypeOf(rec.sensor_id) == RecordType::RobotLaser;
                rec.kind = CarmenRecord::Kind::Scan;
                rec.scan = std::make_shared<ScanData>();
                f.Int();                                 /* laser type */
                const double start_angle = f.Real();
                f.Real();                                /* field of view */
                const double angular_resolution = f.Real();
                const double max_range = f.Real();
                f.Real();                                /* accuracy */
                f.Int();                                 /* remission mode */
                const int n = f.Int();
                f.Ranges(n, rec.scan->ranges);
                if (robot) {
                    const Pose2D laser_pose = f.Pose();
                    rec.odom_pose = f.Pose();
                    rec.velocity.x = f.Real();
                    rec.velocity.theta = f.Real();
                    f.Real(); f.Real(); f.Real();        /* safety distances, turn axis */
                    rec.scan->relative_sensor_pose = InverseCompound(rec.odom_pose, laser_pose);
                } else {
                    const int n_remissions = f.Int();
                    for (int i = 0; i < n_remissions; ++i) f.Real();
                }
                rec.time_stamp = f.Real();
                rec.scan->min_range = 0.0;
                rec.scan->max_range = max_range;
                rec.min_angle = start_angle;
                rec.max_angle = start_angle + angular_resolution * static_cast<double>(n - 1);
                EvenAngles(start_angle, angular_resolution, n, rec.scan->angles);
                records.push_back(std::move(rec));
                break;
            }
            case RecordType::OldLaser:                   /* :320-395 */
            case RecordType::OldOtherLaser: {            /* :398-460 */
                const bool with_poses = TypeOf(rec.sensor_id) == RecordType::OldLaser;
                rec.kind = CarmenRecord::Kind::Scan;
                rec.scan = std::make_shared<ScanData>();
                const int n = f.Int();
                f.Ranges(n, rec.scan->ranges);
                if (with_poses) {
                    const Pose2D laser_pose = f.Pose();
                    rec.odom_pose = f.Pose();
                    rec.time_stamp = f.Real();
                    rec.scan->relative_sensor_pose = InverseCompound(rec.odom_pose, laser_pose);
                }                                        /* LASER3/4: no header is read, the time stamp stays 0 */
                const OldGeometry g = OldLaserGeometry(mParams, n);
                rec.scan->min_range = g.min_range;
                rec.scan->max_range = g.max_range;
                rec.min_angle = g.min_angle;
                rec.max_angle = g.max_angle;
                EvenAngles(g.min_angle, g.increment, n, rec.scan->angles);
                records.push_back(std::move(rec));
                break;
            }
            default:
                break;
        }
    }
    return true;
}

bool CarmenLogReader::LoadFile(const std::string& path, std::vector<CarmenRecord>& records)
{
    std::ifstream in(path);
    if (!in) {
        records.clear();
        return false;
    }
    return Load(in, records);
}

/* ---- writer ------------------------------------------------------------------------ */

namespace {
std::string G17(double v)
{
    char buf[40];
    std::snprintf(buf, sizeof buf, "%.17g", v);
    return buf;
}
}

void CarmenLogWriter::Tail(double time_stamp)
{
    mOut << ' ' << G17(time_stamp) << ' ' << mHost << ' ' << G17(time_stamp) << '\n';
}

void CarmenLogWriter::Param(const std::string& name, const std::string& value)
{
    mOut << "PARAM " << name << ' ' << value << '\n';
}

void CarmenLogWriter::Odom(const Pose2D& pose, double tv, double rv, double time_stamp)
{
    mOut << "ODOM " << G17(pose.x) << ' ' << G17(pose.y) << ' ' << G17(pose.theta) << ' ' << G17(tv) << ' '
         << G17(rv) << " 0";
    Tail(time_stamp);
}

void CarmenLogWriter::RobotLaser(const std::string& sensor_id, double start_angle, double angular_resolution,
                                 double max_range, const std::vector<double>& ranges, const Pose2D& laser_pose,
                                 const Pose2D& robot_pose, double time_stamp)
{
    const double fov = angular_resolution * static_cast<double>(ranges.size());
    mOut << sensor_id << " 0 " << G17(start_angle) << ' ' << G17(fov) << ' ' << G17(angular_resolution) << ' '
         << G17(max_range) << " 0.01 0 " << ranges.size();
    for (double r : ranges) mOut << ' ' << G17(r);
    mOut << ' ' << G17(laser_pose.x) << ' ' << G17(laser_pose.y) << ' ' << G17(laser_pose.theta)
         << ' ' << G17(robot_pose.x) << ' ' << G17(robot_pose.y) << ' ' << G17(robot_pose.theta)
         << " 0 0 0 0 0";
    Tail(time_stamp);
}

void CarmenLogWriter::OldLaser(const std::string& sensor_id, const std::vector<double>& ranges,
                               const Pose2D& laser_pose, const Pose2D& robot_pose, double time_stamp)
{
    mOut << sensor_id << ' ' << ranges.size();
    for (double r : ranges) mOut << ' ' << G17(r);
    mOut << ' ' << G17(laser_pose.x) << ' ' << G17(laser_pose.y) << ' ' << G17(laser_pose.theta)
         << ' ' << G17(robot_pose.x) << ' ' << G17(robot_pose.y) << ' ' << G17(robot_pose.theta);
    Tail(time_stamp);
}

/* ---- metrics file -------------------------------------------------------------------- */

namespace {
/* a JSON string the way boost::property_tree::write_json escapes it */
std::string Quoted(const std::string& s)
{
    std::string out = "\"";
    for (unsigned char c : s) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '/': out += "\\/"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (c < 0x20) {
                    char buf[8];
                    std::snprintf(buf, sizeof buf, "\\u%04X", c);
                    out += buf;
                } else {
                    out += static_cast<char>(c);
                }
        }
    }
    return out + "\"";
}

/* The reference's value sequences are typed: <float> for the ids below, <int> / <uint64_t> for every other
 * one (times in microseconds, counts, sizes); Observe converts the value to that type and VecToString prints
 * floats with six decimals and integers as integers (metric.hpp:42-59). */
bool IsRealValued(const std::string& id)
{
    static const char* const kReal[] = {
        "IntervalAngle", "IntervalTime", "IntervalTravelDist", "LocalMapIntervalTravelDist", "AccumTravelDist",
        "NodeDist", "FinalError", "InitialError", "CostValue", "DiffRotation", "DiffTranslation", "FinalCost",
        "InitialCost", "ScoreValue", "StepSizeTheta", "StepSizeX", "StepSizeY" };
    const std::size_t dot = id.rfind('.');
    const std::string last = dot == std::string::npos ? id : id.substr(dot + 1);
    for (const char* name : kReal)
        if (last == name) return true;
    return false;
}
}

std::string MetricValuesToString(const std::string& id, const std::vector<double>& values)
{
    const bool real = IsRealValued(id);
    std::string joined;
    char buf[352];
    for (std::size_t i = 0; i < values.size(); ++i) {
        if (real)
            std::snprintf(buf, sizeof buf, "%.6f", static_cast<double>(static_cast<float>(values[i])));
        else
            std::snprintf(buf, sizeof buf, "%lld", static_cast<long long>(values[i]));
        if (i) joined += ' ';
        joined += buf;
    }
    return joined;
}

void WriteMetricsJson(std::ostream& out, const MetricRecorder& metrics)
{
    /* metric.cpp:460-496: five families; this package reports value sequences only (metrics.hpp), an empty
     * family is an empty property tree, which write_json prints as "" */
    out << "{\n";
    for (const char* family : { "Counters", "Gauges", "Distributions", "Histograms" })
        out << "    " << Quoted(family) << ": \"\",\n";
    const auto& values = metrics.Values();
    if (values.empty()) {
        out << "    \"ValueSequences\": \"\"\n}\n";
        return;
    }
    out << "    \"ValueSequences\": {\n";
    std::size_t k = 0;
    for (const auto& kv : values) {
        /* metric.hpp:611-621 */
        const std::string joined = MetricValuesToString(kv.first, kv.second);
        out << "        " << Quoted(kv.first) << ": {\n"
            << "            \"NumOfSamples\": \"" << kv.second.size() << "\",\n"
            << "            \"Values\": " << Quoted(joined) << "\n"
            << "        }" << (++k < values.size() ? "," : "") << "\n";
    }
    out << "    }\n}\n";
}

bool SaveMetrics(const std::string& output_path, const MetricRecorder& metrics)
{
    std::ofstream out(output_path + ".metric.json");
    if (!out) return false;
    WriteMetricsJson(out, metrics);
    return static_cast<bool>(out);
}

} /* namespace csm_host */
