// compat.hpp -- the slice of Eigen and PCL that the FastVGICPCuda / LsqRegistration interface touches, for builds where
// neither library exists (this image has no Eigen, PCL, Boost or FLANN).  When the real headers are available define
// FAST_GICP_B200_USE_SYSTEM_PCL and include them before this file; the class templates in lsq_registration.hpp and
// fast_vgicp_cuda.hpp only use the members declared here, which are spelled exactly like the real ones
// (pcl::PointXYZ::getVector3fMap is replaced by direct x/y/z access, the one deliberate difference).
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#ifndef FAST_GICP_B200_USE_SYSTEM_PCL

namespace Eigen {

// fixed-size, column-major, like Eigen::Matrix<Scalar, Rows, Cols>
template <typename Scalar, int Rows, int Cols>
struct Matrix {
  std::array<Scalar, Rows * Cols> v{};
  Scalar& operator()(int r, int c) { return v[c * Rows + r]; }
  const Scalar& operator()(int r, int c) const { return v[c * Rows + r]; }
  Scalar& operator[](int i) { return v[i]; }
  const Scalar& operator[](int i) const { return v[i]; }
  Scalar* data() { return v.data(); }
  const Scalar* data() const { return v.data(); }
  void setZero() { v.fill(Scalar(0)); }
  void setIdentity() {
    setZero();
    for (int i = 0; i < (Rows < Cols ? Rows : Cols); i++) (*this)(i, i) = Scalar(1);
  }
  static Matrix Identity() {
    Matrix m;
    m.setIdentity();
    return m;
  }
  static Matrix Zero() { return Matrix(); }
  template <typename T>
  Matrix<T, Rows, Cols> cast() const {
    Matrix<T, Rows, Cols> o;
    for (int i = 0; i < Rows * Cols; i++) o.v[i] = static_cast<T>(v[i]);
    return o;
  }
  Matrix operator*(const Matrix& b) const {  // square product (used for 4x4 poses)
    static_assert(Rows == Cols, "square only");
    Matrix o;
    for (int c = 0; c < Cols; c++)
      for (int r = 0; r < Rows; r++) {
        Scalar s = 0;
        for (int k = 0; k < Cols; k++) s += (*this)(r, k) * b(k, c);
        o(r, c) = s;
      }
    return o;
  }
};
using Matrix4f = Matrix<float, 4, 4>;
using Matrix4d = Matrix<double, 4, 4>;

// rigid inverse of a 4x4 isometry image
template <typename Scalar>
Matrix<Scalar, 4, 4> isometry_inverse(const Matrix<Scalar, 4, 4>& T) {
  Matrix<Scalar, 4, 4> o = Matrix<Scalar, 4, 4>::Identity();
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o(r, c) = T(c, r);
  for (int r = 0; r < 3; r++) o(r, 3) = -(o(r, 0) * T(0, 3) + o(r, 1) * T(1, 3) + o(r, 2) * T(2, 3));
  return o;
}

}  // namespace Eigen

namespace pcl {

struct PointXYZ {
  float x = 0, y = 0, z = 0, pad = 1.0f;  // 16 bytes like PCL's SSE-aligned point
  PointXYZ() = default;
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
struct PointXYZI {
  float x = 0, y = 0, z = 0, pad = 1.0f;
  float intensity = 0, pad2[3] = {0, 0, 0};  // 32 bytes
};

template <typename PointT>
class PointCloud {
public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void resize(std::size_t n) { points.resize(n); }
  void clear() { points.clear(); }
  void push_back(const PointT& p) { points.push_back(p); }
  PointT& at(std::size_t i) { return points.at(i); }
  const PointT& at(std::size_t i) const { return points.at(i); }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  typename std::vector<PointT>::iterator begin() { return points.begin(); }
  typename std::vector<PointT>::iterator end() { return points.end(); }
  typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
  typename std::vector<PointT>::const_iterator end() const { return points.end(); }
};

template <typename T, typename... Args>
std::shared_ptr<T> make_shared(Args&&... args) {
  return std::make_shared<T>(std::forward<Args>(args)...);
}

// pcl::Registration<PointSource, PointTarget, Scalar>: the members the reference's classes use
// (lsq_registration.hpp:38-43, fast_vgicp_cuda.hpp:48-50) and the public calls of apps/tests/bindings.
template <typename PointSource, typename PointTarget, typename Scalar = float>
class Registration {
public:
  using Matrix4 = Eigen::Matrix<Scalar, 4, 4>;
  using PointCloudSource = PointCloud<PointSource>;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = PointCloud<PointTarget>;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  using Ptr = std::shared_ptr<Registration<PointSource, PointTarget, Scalar>>;

  Registration() { final_transformation_.setIdentity(); }
  virtual ~Registration() {}

  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { input_ = cloud; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) { target_ = cloud; }
  PointCloudSourceConstPtr getInputSource() const { return input_; }
  PointCloudTargetConstPtr getInputTarget() const { return target_; }

  void setMaximumIterations(int n) { max_iterations_ = n; }
  int getMaximumIterations() const { return max_iterations_; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setMaxCorrespondenceDistance(double d) { corr_dist_threshold_ = d; }
  Matrix4 getFinalTransformation() const { return final_transformation_; }
  bool hasConverged() const { return converged_; }
  const std::string& getClassName() const { return reg_name_; }

  // pcl::Registration::align: initCompute, output := input, reset state, computeTransformation(output, guess)
  void align(PointCloudSource& output) { align(output, Matrix4::Identity()); }
  void align(PointCloudSource& output, const Matrix4& guess) {
    if (!input_ || !target_) return;  // PCL: initCompute() fails and align returns
    output = *input_;
    converged_ = false;
    final_transformation_.setIdentity();
    computeTransformation(output, guess);
  }

protected:
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;

  std::string reg_name_ = "Registration";
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  int nr_iterations_ = 0;
  int max_iterations_ = 10;
  Matrix4 final_transformation_;
  double transformation_epsilon_ = 0.0;
  double corr_dist_threshold_ = std::sqrt(std::numeric_limits<double>::max());
  bool converged_ = false;
};

}  // namespace pcl

#endif  // FAST_GICP_B200_USE_SYSTEM_PCL
