// Exercises the C++ drop-in (mono_vo::FeatureProcessor + mono_vo::gpu::*) the way the reference's
// Frame / Initializer / Tracker call it.  Reads two raw 8-bit frames, prints one line of counters.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "mono_vo/feature_processor.hpp"
#include "mono_vo/gpu_cv.hpp"

static cv::Mat load(const char* path, int w, int h) {
  cv::Mat m(h, w, CV_8UC1);
  FILE* f = fopen(path, "rb");
  if (!f || fread(m.data, 1, (size_t)w * h, f) != (size_t)w * h) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  fclose(f);
  return m;
}

int main(int argc, char** argv) {
  if (argc < 6) { fprintf(stderr, "usage: test_adapter f0.raw f1.raw w h nfeatures\n"); return 2; }
  const int w = atoi(argv[3]), h = atoi(argv[4]), nf = atoi(argv[5]);
  cv::Mat f0 = load(argv[1], w, h), f1 = load(argv[2], w, h);
  auto fp = std::make_shared<mono_vo::FeatureProcessor>(nf);
  std::vector<cv::KeyPoint> k0, k1;
  cv::Mat d0, d1;
  fp->detect_and_compute(f0, k0, d0);          // Frame::extract_observations, src/frame.cpp:12
  fp->detect_and_compute(f1, k1, d1);
  std::vector<cv::KeyPoint> kd = fp->detect(f0);
  std::vector<cv::DMatch> m = fp->find_matches(d0, d1, 0.7);   // src/initializer.cpp:187
  std::vector<cv::Point2f> p0, p1;
  for (auto& k : k0) p0.push_back(k.pt);
  std::vector<unsigned char> status, mh, mf, me;
  std::vector<float> err;
  mvo_ctx* c = fp->context(w, h);
  mono_vo::gpu::calcOpticalFlowPyrLK(c, f0, f1, p0, p1, status, err);   // src/tracker.cpp:68
  std::vector<cv::Point2f> a, b;
  for (size_t i = 0; i < p0.size(); ++i) if (status[i] && err[i] < 30.0f) { a.push_back(p0[i]); b.push_back(p1[i]); }
  cv::Mat K(3, 3, CV_64F);
  K.at<double>(0, 0) = K.at<double>(1, 1) = 718.856 * w / 1241.0;
  K.at<double>(0, 2) = (w - 1) / 2.0; K.at<double>(1, 2) = (h - 1) / 2.0; K.at<double>(2, 2) = 1;
  cv::Mat H = mono_vo::gpu::findHomography(c, a, b, cv::RANSAC, 1.0, mh);
  cv::Mat F = mono_vo::gpu::findFundamentalMat(c, a, b, cv::FM_RANSAC, 1.0, 0.99, mf);
  cv::Mat E = mono_vo::gpu::findEssentialMat(c, a, b, K, cv::RANSAC, 0.99, 1.0, me);
  cv::Mat R, t;
  int good = mono_vo::gpu::recoverPose(c, E, a, b, K, R, t, me);
  int sh = 0, sf = 0;
  for (auto v : mh) sh += v != 0;
  for (auto v : mf) sf += v != 0;
  // Tracker::update, src/tracker.cpp:298-316: landmarks + their observations -> pose
  int pnp_inl = -1;
  double rt[6] = {0, 0, 0, 0, 0, 0};
  if (argc >= 9) {
    const int np = atoi(argv[8]);
    std::vector<cv::Point3f> obj(np);
    std::vector<cv::Point2f> img(np);
    FILE* fo = fopen(argv[6], "rb");
    FILE* fi = fopen(argv[7], "rb");
    if (!fo || !fi || fread(obj.data(), 12, np, fo) != (size_t)np || fread(img.data(), 8, np, fi) != (size_t)np) return 2;
    fclose(fo);
    fclose(fi);
    cv::Mat rvec, tvec, Rm, dist(1, 5, CV_64F);
    std::vector<int> inl;
    cv::Mat Kp(3, 3, CV_64F);
    Kp.at<double>(0, 0) = Kp.at<double>(1, 1) = 718.856; Kp.at<double>(0, 2) = 620.5; Kp.at<double>(1, 2) = 188.0; Kp.at<double>(2, 2) = 1;
    if (mono_vo::gpu::solvePnPRansac(c, obj, img, Kp, dist, rvec, tvec, false, 100, 8.0, 0.99, inl)) {
      pnp_inl = (int)inl.size();
      mono_vo::gpu::Rodrigues(rvec, Rm);
      for (int i = 0; i < 3; ++i) { rt[i] = rvec.at<double>(i, 0); rt[3 + i] = tvec.at<double>(i, 0); }
    }
  }
  printf("pnp %d %.12g %.12g %.12g %.12g %.12g %.12g ", pnp_inl, rt[0], rt[1], rt[2], rt[3], rt[4], rt[5]);
  printf("kps %zu %zu detect %zu matches %zu tracked %zu score_h %d score_f %d good %d desc %dx%d first %.3f %.3f %d\n",
         k0.size(), k1.size(), kd.size(), m.size(), a.size(), sh, sf, good, d0.rows, d0.cols, k0[0].pt.x, k0[0].pt.y,
         (int)d0.at<unsigned char>(0, 0));
  return 0;
}
