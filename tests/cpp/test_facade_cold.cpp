// Cold helpers of the OpenCV-API facade (ros2_mono_vo_b200/cpp/facade): the small cv::Mat / cv::Affine3d algebra the
// reference's Initializer / Tracker use around the hot calls.  Host only -- no GPU call is made here.
#include <cmath>
#include <cstdio>
#include <sstream>
#include <opencv2/opencv.hpp>

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)
static bool near(double a, double b, double eps = 1e-12) { return std::fabs(a - b) <= eps * (1 + std::fabs(b)); }

int main() {
  // K * [I | 0] and K * [R | t]  (src/initializer.cpp:117-122)
  cv::Mat K = (cv::Mat_<double>(3, 3) << 700, 0, 320, 0, 710, 240, 0, 0, 1);
  cv::Mat P0 = K * cv::Mat::eye(3, 4, CV_64F);
  CHECK(P0.rows == 3 && P0.cols == 4 && near(P0.at<double>(0, 0), 700) && near(P0.at<double>(1, 2), 240) && near(P0.at<double>(2, 3), 0));
  const double a = 0.3;
  cv::Mat R = (cv::Mat_<double>(3, 3) << std::cos(a), -std::sin(a), 0, std::sin(a), std::cos(a), 0, 0, 0, 1);
  cv::Mat t = (cv::Mat_<double>(3, 1) << 0.1, -0.2, 0.9);
  cv::Mat Rt;
  cv::hconcat(R, t, Rt);
  CHECK(Rt.rows == 3 && Rt.cols == 4 && near(Rt.at<double>(1, 3), -0.2) && near(Rt.at<double>(0, 1), -std::sin(a)));
  cv::Mat P1 = K * Rt;
  CHECK(near(P1.at<double>(0, 3), 700 * 0.1 + 320 * 0.9));
  // R * p + t  (the chirality lambda, src/initializer.cpp:142-146)
  cv::Mat p = (cv::Mat_<double>(3, 1) << 1.f, 2.f, 3.f);
  cv::Mat q = R * p + t;
  CHECK(near(q.at<double>(0, 0), std::cos(a) - 2 * std::sin(a) + 0.1) && near(q.at<double>(2, 0), 3.9));
  // transpose, ROI, row views share storage, clone does not
  cv::Mat Pt = P1.t();
  CHECK(Pt.rows == 4 && Pt.cols == 3 && near(Pt.at<double>(3, 0), P1.at<double>(0, 3)));
  cv::Mat D(4, 32, CV_8UC1);
  cv::Mat r2 = D.row(2);
  r2.at<uchar>(0, 5) = 77;
  CHECK(D.at<uchar>(2, 5) == 77);
  cv::Mat C = D.clone();
  C.at<uchar>(2, 5) = 1;
  CHECK(D.at<uchar>(2, 5) == 77);
  cv::Mat one(1, 32, CV_8UC1);
  one.at<uchar>(0, 31) = 9;
  one.copyTo(D.row(3));   // Frame::get_descriptors, src/frame.cpp:59-61
  CHECK(D.at<uchar>(3, 31) == 9);
  // Affine3d: construction from cv::Mat R / t, inverse, composition, point transform, Mat(matrix)(Rect)
  cv::Affine3d T(R, t);
  cv::Affine3d Ti = T.inv();
  cv::Affine3d I = T * Ti;
  CHECK(cv::norm(I.matrix - cv::Affine3d().matrix) < 1e-14);
  cv::Point3f x(1.f, 2.f, 3.f);
  cv::Point3f y = T * x;
  CHECK(near(y.x, q.at<double>(0, 0), 1e-6) && near(y.z, 3.9, 1e-6));
  cv::Vec3d tr = T.translation();
  CHECK(near(cv::norm(tr), std::sqrt(0.01 + 0.04 + 0.81)));
  cv::Matx33d Rm = T.rotation();
  CHECK(near(Rm(0, 1), -std::sin(a)));
  cv::Mat ext = cv::Mat(Ti.matrix)(cv::Rect(0, 0, 4, 3));   // src/tracker.cpp:143-146
  CHECK(ext.rows == 3 && ext.cols == 4);
  cv::Mat Pc = K * ext;
  CHECK(Pc.rows == 3 && Pc.cols == 4 && near(Pc.at<double>(2, 3), Ti.matrix(2, 3)));
  std::stringstream ss;
  ss << T.matrix;
  CHECK(ss.str().find("0.9") != std::string::npos);
  // convertPointsFromHomogeneous on a transposed 4 x N float matrix, both output forms (src/initializer.cpp:131, tracker.cpp:152)
  cv::Mat X4(4, 3, CV_32F);
  const float cols[3][4] = {{2, 4, 6, 2}, {1, 1, 1, 0}, {-3, 6, 9, -3}};
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 4; ++k) X4.at<float>(k, i) = cols[i][k];
  std::vector<cv::Point3f> pts;
  cv::convertPointsFromHomogeneous(X4.t(), pts);
  CHECK(pts.size() == 3 && pts[0].x == 1.f && pts[0].z == 3.f && pts[1].x == 1.f /* w == 0 -> scale 1 */ && pts[2].y == -2.f);
  cv::Mat pm;
  cv::convertPointsFromHomogeneous(X4.t(), pm);
  CHECK(pm.rows == 3 && pm.cols == 1 && pm.at<cv::Point3f>(2).z == -3.f);
  // countNonZero, keypoint grid (Initializer::good_keypoint_distribution, src/initializer.cpp:57-66)
  std::vector<uchar> mask = {1, 0, 2, 0, 1};
  CHECK(cv::countNonZero(mask) == 3);
  cv::Mat grid = cv::Mat::zeros(376 / 50, 1241 / 50, CV_8U);
  CHECK(grid.rows == 7 && grid.cols == 24 && !grid.at<uchar>(6, 23));
  grid.at<uchar>(6, 23) = 1;
  CHECK(cv::countNonZero(grid) == 1);
  // Rodrigues is host arithmetic behind the ABI (no device needed)
  cv::Mat rvec = (cv::Mat_<double>(3, 1) << 0, 0, a), Rr;
  cv::Rodrigues(rvec, Rr);
  CHECK(near(Rr.at<double>(0, 0), std::cos(a)) && near(Rr.at<double>(1, 0), std::sin(a)));
  // degenerate inputs return OpenCV's empties without touching the GPU
  std::vector<cv::Point2f> few(3), few2(3);
  std::vector<uchar> m;
  CHECK(cv::findHomography(few, few2, cv::RANSAC, 1.0, m).empty() && m.size() == 3);
  CHECK(cv::findFundamentalMat(few, few2, cv::FM_RANSAC, 1.0, 0.99, m).empty());
  CHECK(cv::findEssentialMat(few, few2, K, cv::RANSAC, 0.99, 1.0, m).empty());
  std::vector<std::vector<cv::DMatch>> knn;
  cv::BFMatcher(cv::NORM_HAMMING).knnMatch(cv::Mat(), cv::Mat(), knn, 2);
  CHECK(knn.empty());
  std::printf(fails ? "FAILED %d\n" : "ok\n", fails);
  return fails ? 1 : 0;
}
