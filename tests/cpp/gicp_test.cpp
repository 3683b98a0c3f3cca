// The reference's alignment test (src/test/gicp_test.cpp:147-201) re-expressed for the VGICP_CUDA method on the C++
// mirror classes: forward, backward, swap+setSource, swap+setTarget, each against data/relative.txt within 0.05 m / 1 deg
// plus hasConverged().  No gtest in this image: plain main, exit code = number of failed expectations.
//   usage: gicp_test target.bin n_target source.bin n_source relative.txt     (clouds: packed float32 xyz)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include <fast_gicp_b200/fast_vgicp_cuda.hpp>

using Cloud = pcl::PointCloud<pcl::PointXYZ>;
using Reg = fast_gicp::FastVGICPCuda<pcl::PointXYZ, pcl::PointXYZ>;

static Cloud::Ptr load(const char* path, size_t n) {
  auto c = pcl::make_shared<Cloud>();
  std::ifstream f(path, std::ios::binary);
  std::vector<float> raw(n * 3);
  f.read(reinterpret_cast<char*>(raw.data()), raw.size() * sizeof(float));
  if (!f) { std::fprintf(stderr, "cannot read %s\n", path); std::exit(100); }
  c->resize(n);
  for (size_t i = 0; i < n; i++) { c->at(i).x = raw[3 * i]; c->at(i).y = raw[3 * i + 1]; c->at(i).z = raw[3 * i + 2]; }
  return c;
}

static int failures = 0;
#define EXPECT(cond, what) do { if (!(cond)) { std::printf("FAILED: %s (%s)\n", what, #cond); failures++; } } while (0)

static void pose_error(const Eigen::Matrix4f& gt, const Eigen::Matrix4f& est, double& t_err, double& r_err) {  // gicp_test.cpp:75-80
  Eigen::Matrix4f d = Eigen::isometry_inverse(gt) * est;
  t_err = std::sqrt(d(0, 3) * d(0, 3) + d(1, 3) * d(1, 3) + d(2, 3) * d(2, 3));
  double c = (d(0, 0) + d(1, 1) + d(2, 2) - 1.0) / 2.0;
  r_err = std::acos(std::fmin(1.0, std::fmax(-1.0, c)));
}

int main(int argc, char** argv) {
  if (argc < 6) { std::fprintf(stderr, "usage: %s target.bin n source.bin n relative.txt\n", argv[0]); return 100; }
  Cloud::ConstPtr target = load(argv[1], std::strtoul(argv[2], nullptr, 10));
  Cloud::ConstPtr source = load(argv[3], std::strtoul(argv[4], nullptr, 10));
  Eigen::Matrix4f gt;
  {
    std::ifstream f(argv[5]);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) f >> gt(i, j);
  }
  const double t_tol = 0.05, r_tol = 1.0 * M_PI / 180.0;
  double te, re;
  Cloud aligned;

  {  // forward test
    Reg reg;
    reg.setInputTarget(target);
    reg.setInputSource(source);
    reg.align(aligned);
    pose_error(gt, reg.getFinalTransformation(), te, re);
    EXPECT(te < t_tol && re < r_tol && reg.hasConverged(), "FORWARD TEST");
    EXPECT(aligned.size() == source->size(), "aligned cloud size");
    std::printf("forward: t_err %.4f m r_err %.4f deg H(0,0)=%.3g\n", te, re * 180 / M_PI, reg.getFinalHessian()(0, 0));
    // backward test on the same object
    reg.setInputTarget(source);
    reg.setInputSource(target);
    reg.align(aligned);
    pose_error(gt, Eigen::isometry_inverse(reg.getFinalTransformation()), te, re);
    EXPECT(te < t_tol && re < r_tol && reg.hasConverged(), "BACKWARD TEST");
  }
  {  // swap and set source
    Reg reg;
    reg.setInputSource(target);
    reg.swapSourceAndTarget();
    reg.setInputSource(source);
    reg.align(aligned);
    pose_error(gt, reg.getFinalTransformation(), te, re);
    EXPECT(te < t_tol && re < r_tol && reg.hasConverged(), "SWAP AND SET SOURCE TEST");
  }
  {  // swap and set target
    Reg reg;
    reg.setInputTarget(source);
    reg.swapSourceAndTarget();
    reg.setInputTarget(target);
    reg.align(aligned);
    pose_error(gt, reg.getFinalTransformation(), te, re);
    EXPECT(te < t_tol && re < r_tol && reg.hasConverged(), "SWAP AND SET TARGET TEST");
  }
  {  // DIRECT27 + Gauss-Newton through the same interface
    Reg reg;
    reg.setNeighborSearchMethod(fast_gicp::NeighborSearchMethod::DIRECT27);
    reg.setOptimizerType(fast_gicp::LSQ_OPTIMIZER_TYPE::GaussNewton);
    reg.setInputTarget(target);
    reg.setInputSource(source);
    reg.align(aligned);
    pose_error(gt, reg.getFinalTransformation(), te, re);
    EXPECT(te < t_tol && re < r_tol && reg.hasConverged(), "DIRECT27 GAUSS-NEWTON");
  }
  std::printf("%s (%d failures)\n", failures ? "FAILED" : "PASSED", failures);
  return failures;
}
