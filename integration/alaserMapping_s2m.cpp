// integration/alaserMapping_s2m.cpp -- ROS1 node `alaserMapping` with the body of process() replaced by
// libs2m.  NOT compiled in this repository: the image has no ROS / PCL.  It is the
// source a maintainer adds to the catkin package in place of src/laserMapping.cpp
// (CMakeLists.txt:52-53: add_executable(alaserMapping ...); link `s2m` instead of Ceres).
//
// What stays as in the reference node (graph contract):
//   subscriptions  /laser_cloud_corner_last, /laser_cloud_surf_last, /velodyne_cloud_3, /laser_odom_to_init
//                  (laserMapping.cpp:921-927, queue 100)
//   publications   /laser_cloud_surround, /laser_cloud_map, /velodyne_cloud_registered,
//                  /velodyne_cloud_registered_local, /aft_mapped_to_init, /aft_mapped_to_init_high_frec,
//                  /aft_mapped_path, TF /camera_init -> /aft_mapped (:929-940, :888-899)
//   parameters     mapping_line_resolution (0.4), mapping_plane_resolution (0.8) (:913-919)
//   message rules  odom / surf / full older than the front corner stamp are dropped; a set is used only
//                  when all four stamps are equal; after taking a set the corner queue is flushed
//                  (:236-304); every 5th frame the surround cloud, every 20th the whole map (:807-837)
// What changes: rows A ... W of process() (:310-802) are one call, s2m_register.
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include <nav_msgs/Odometry.h>
#include <nav_msgs/Path.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl_conversions/pcl_conversions.h>
#include <ros/ros.h>
#include <sensor_msgs/PointCloud2.h>
#include <tf/transform_broadcaster.h>

#include "s2m.h"

namespace {

template <typename Msg>
class Inbox {  // one subscription's queue; callbacks run on the spin thread, the worker drains it
 public:
  void put(const Msg& m) { std::lock_guard<std::mutex> g(mu_); q_.push_back(m); }
  bool empty() { std::lock_guard<std::mutex> g(mu_); return q_.empty(); }
  double front_stamp() { std::lock_guard<std::mutex> g(mu_); return q_.front()->header.stamp.toSec(); }
  Msg take() { std::lock_guard<std::mutex> g(mu_); Msg m = q_.front(); q_.pop_front(); return m; }
  void drop_older_than(double t) {
    std::lock_guard<std::mutex> g(mu_);
    while (!q_.empty() && q_.front()->header.stamp.toSec() < t) q_.pop_front();
  }
  void clear() { std::lock_guard<std::mutex> g(mu_); q_.clear(); }

 private:
  std::mutex mu_;
  std::deque<Msg> q_;
};

using CloudMsg = sensor_msgs::PointCloud2ConstPtr;
using Cloud = pcl::PointCloud<pcl::PointXYZI>;

std::vector<float> pack(const CloudMsg& msg, int* n) {  // PointXYZI on the wire -> packed xyzi
  Cloud c;
  pcl::fromROSMsg(*msg, c);
  std::vector<float> out(4 * c.size());
  for (size_t i = 0; i < c.size(); ++i) {
    out[4 * i] = c[i].x; out[4 * i + 1] = c[i].y; out[4 * i + 2] = c[i].z; out[4 * i + 3] = c[i].intensity;
  }
  *n = (int)c.size();
  return out;
}
sensor_msgs::PointCloud2 unpack(const std::vector<float>& xyzi, int n, const ros::Time& stamp, const char* frame) {
  Cloud c;
  c.resize(n);
  for (int i = 0; i < n; ++i) {
    c[i].x = xyzi[4 * i]; c[i].y = xyzi[4 * i + 1]; c[i].z = xyzi[4 * i + 2]; c[i].intensity = xyzi[4 * i + 3];
  }
  sensor_msgs::PointCloud2 msg;
  pcl::toROSMsg(c, msg);
  msg.header.stamp = stamp;
  msg.header.frame_id = frame;
  return msg;
}

class MappingNode {
 public:
  explicit MappingNode(ros::NodeHandle& nh) {
    float line_res = 0.4f, plane_res = 0.8f;
    nh.param<float>("mapping_line_resolution", line_res, 0.4f);
    nh.param<float>("mapping_plane_resolution", plane_res, 0.8f);
    s2m_params p;
    s2m_default_params(&p);
    p.line_res = line_res;  // the floats setLeafSize received (laserMapping.cpp:918-919)
    p.plane_res = plane_res;
    if (s2m_create(&p, &s2m_) != S2M_OK) { ROS_FATAL("s2m_create failed (no CUDA device?)"); ros::shutdown(); return; }
    sub_corner_ = nh.subscribe<sensor_msgs::PointCloud2>("/laser_cloud_corner_last", 100, [this](const CloudMsg& m) { corner_.put(m); });
    sub_surf_ = nh.subscribe<sensor_msgs::PointCloud2>("/laser_cloud_surf_last", 100, [this](const CloudMsg& m) { surf_.put(m); });
    sub_full_ = nh.subscribe<sensor_msgs::PointCloud2>("/velodyne_cloud_3", 100, [this](const CloudMsg& m) { full_.put(m); });
    sub_odom_ = nh.subscribe<nav_msgs::Odometry>("/laser_odom_to_init", 100, [this](const nav_msgs::Odometry::ConstPtr& m) { on_odometry(m); });
    pub_surround_ = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_surround", 100);
    pub_map_ = nh.advertise<sensor_msgs::PointCloud2>("/laser_cloud_map", 100);
    pub_registered_ = nh.advertise<sensor_msgs::PointCloud2>("/velodyne_cloud_registered", 100);
    pub_registered_local_ = nh.advertise<sensor_msgs::PointCloud2>("/velodyne_cloud_registered_local", 100);
    pub_odom_ = nh.advertise<nav_msgs::Odometry>("/aft_mapped_to_init", 100);
    pub_odom_fast_ = nh.advertise<nav_msgs::Odometry>("/aft_mapped_to_init_high_frec", 100);
    pub_path_ = nh.advertise<nav_msgs::Path>("/aft_mapped_path", 100);
    worker_ = std::thread([this] { run(); });
  }
  ~MappingNode() {
    if (worker_.joinable()) worker_.join();
    s2m_destroy(s2m_);
  }

 private:
  // high-rate relay (:198-230): odometry pose composed with the correction the library keeps (row U)
  void on_odometry(const nav_msgs::Odometry::ConstPtr& m) {
    odom_.put(m);
    double qc[4], tc[3];
    {
      std::lock_guard<std::mutex> g(s2m_mu_);
      s2m_get_correction(s2m_, 0, qc, tc);
    }
    const auto& o = m->pose.pose.orientation;
    const auto& t = m->pose.pose.position;
    const double qo[4] = {o.x, o.y, o.z, o.w};
    // q = qc * qo ; p = qc * t + tc   (Hamilton product, x y z w)
    const double q[4] = {qc[3] * qo[0] + qc[0] * qo[3] + qc[1] * qo[2] - qc[2] * qo[1], qc[3] * qo[1] - qc[0] * qo[2] + qc[1] * qo[3] + qc[2] * qo[0],
                         qc[3] * qo[2] + qc[0] * qo[1] - qc[1] * qo[0] + qc[2] * qo[3], qc[3] * qo[3] - qc[0] * qo[0] - qc[1] * qo[1] - qc[2] * qo[2]};
    const double v[3] = {t.x, t.y, t.z};
    const double u[3] = {2 * (qc[1] * v[2] - qc[2] * v[1]), 2 * (qc[2] * v[0] - qc[0] * v[2]), 2 * (qc[0] * v[1] - qc[1] * v[0])};
    const double p[3] = {v[0] + qc[3] * u[0] + qc[1] * u[2] - qc[2] * u[1] + tc[0], v[1] + qc[3] * u[1] + qc[2] * u[0] - qc[0] * u[2] + tc[1],
                         v[2] + qc[3] * u[2] + qc[0] * u[1] - qc[1] * u[0] + tc[2]};
    pub_odom_fast_.publish(make_odometry(q, p, m->header.stamp));
  }
  static nav_msgs::Odometry make_odometry(const double q[4], const double p[3], const ros::Time& stamp) {
    nav_msgs::Odometry o;
    o.header.frame_id = "/camera_init";
    o.child_frame_id = "/aft_mapped";
    o.header.stamp = stamp;
    o.pose.pose.orientation.x = q[0]; o.pose.pose.orientation.y = q[1]; o.pose.pose.orientation.z = q[2]; o.pose.pose.orientation.w = q[3];
    o.pose.pose.position.x = p[0]; o.pose.pose.position.y = p[1]; o.pose.pose.position.z = p[2];
    return o;
  }
  void run() {
    ros::Rate idle(500);
    long frame = 0;
    while (ros::ok()) {
      if (corner_.empty() || surf_.empty() || full_.empty() || odom_.empty()) { idle.sleep(); continue; }
      const double t0 = corner_.front_stamp();
      odom_.drop_older_than(t0); surf_.drop_older_than(t0); full_.drop_older_than(t0);
      if (surf_.empty() || full_.empty() || odom_.empty()) { idle.sleep(); continue; }
      if (surf_.front_stamp() != t0 || full_.front_stamp() != t0 || odom_.front_stamp() != t0) { idle.sleep(); continue; }  // unsync message
      const CloudMsg mc = corner_.take(), ms = surf_.take(), mf = full_.take();
      const nav_msgs::Odometry::ConstPtr mo = odom_.take();
      corner_.clear();  // real time: later corner clouds are dropped (:301-304)
      int nc, ns, nf;
      const std::vector<float> c = pack(mc, &nc), s = pack(ms, &ns), f = pack(mf, &nf);
      const auto& o = mo->pose.pose.orientation;
      const auto& t = mo->pose.pose.position;
      const double q_odom[4] = {o.x, o.y, o.z, o.w}, t_odom[3] = {t.x, t.y, t.z};
      double q_w[4], t_w[3];
      s2m_stats st;
      std::vector<float> registered(f.size());
      int rc;
      {
        std::lock_guard<std::mutex> g(s2m_mu_);
        rc = s2m_register(s2m_, c.data(), nc, s.data(), ns, q_odom, t_odom, q_w, t_w, &st);  // rows A ... W
        if (rc >= 0) s2m_transform_cloud(s2m_, 0, f.data(), nf, registered.data());          // row X (:845-849)
      }
      if (rc == S2M_MAP_TOO_SMALL) ROS_WARN("time Map corner and surf num are not enough");  // :733
      else if (rc < 0) { ROS_ERROR("s2m_register: %s (%s)", s2m_strerror(rc), s2m_last_error(s2m_)); continue; }
      const ros::Time stamp = mo->header.stamp;
      if (frame % 5 == 0) publish_cloud(pub_surround_, stamp, [this](float* out, int cap) { return s2m_get_surround(s2m_, 0, out, cap); });
      if (frame % 20 == 0) {
        std::vector<float> all;
        int n_all = 0;
        for (int cls = 0; cls < 2; ++cls) {
          std::lock_guard<std::mutex> g(s2m_mu_);
          const int n = s2m_map_download(s2m_, 0, cls, nullptr, 0);
          all.resize(4 * (size_t)(n_all + n));
          s2m_map_download(s2m_, 0, cls, all.data() + 4 * (size_t)n_all, n);
          n_all += n;
        }
        pub_map_.publish(unpack(all, n_all, stamp, "/camera_init"));
      }
      pub_registered_local_.publish(unpack(f, nf, stamp, "/camera_init"));   // the untouched full cloud (:839-843)
      pub_registered_.publish(unpack(registered, nf, stamp, "/camera_init"));
      const nav_msgs::Odometry aft = make_odometry(q_w, t_w, stamp);
      pub_odom_.publish(aft);
      geometry_msgs::PoseStamped ps;
      ps.header = aft.header;
      ps.pose = aft.pose.pose;
      path_.header = aft.header;
      path_.poses.push_back(ps);
      pub_path_.publish(path_);
      tf::Transform tr;
      tr.setOrigin(tf::Vector3(t_w[0], t_w[1], t_w[2]));
      tr.setRotation(tf::Quaternion(q_w[0], q_w[1], q_w[2], q_w[3]));
      tf_.sendTransform(tf::StampedTransform(tr, stamp, "/camera_init", "/aft_mapped"));
      ++frame;
    }
  }
  template <typename Getter>
  void publish_cloud(ros::Publisher& pub, const ros::Time& stamp, Getter get) {
    std::vector<float> buf;
    int n;
    {
      std::lock_guard<std::mutex> g(s2m_mu_);
      n = get(nullptr, 0);
      buf.resize(4 * (size_t)std::max(n, 1));
      get(buf.data(), n);
    }
    pub.publish(unpack(buf, n, stamp, "/camera_init"));
  }

  s2m_ctx* s2m_ = nullptr;
  std::mutex s2m_mu_;  // the context is single-caller: the relay on the spin thread only reads the correction
  Inbox<CloudMsg> corner_, surf_, full_;
  Inbox<nav_msgs::Odometry::ConstPtr> odom_;
  ros::Subscriber sub_corner_, sub_surf_, sub_full_, sub_odom_;
  ros::Publisher pub_surround_, pub_map_, pub_registered_, pub_registered_local_, pub_odom_, pub_odom_fast_, pub_path_;
  nav_msgs::Path path_;
  tf::TransformBroadcaster tf_;
  std::thread worker_;
};

}  // namespace

int main(int argc, char** argv) {
  ros::init(argc, argv, "laserMapping");
  ros::NodeHandle nh;
  MappingNode node(nh);
  ros::spin();
  return 0;
}
