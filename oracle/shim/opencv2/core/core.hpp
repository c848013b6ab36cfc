// oracle/shim/opencv2/core/core.hpp -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
//
// OpenCV's C++ headers are not installed in this image.  The reference's hot path uses cv::Mat purely as a
// row-major FP64 container (GCRANSAC.h, scoring_function.h, samplers/*, neighborhood/grid_neighborhood_graph.h
// touch only rows/cols/data/type()/row()/ptr<>()/at<>()).  This header provides exactly that container surface so
// those reference headers compile UNMODIFIED from /root/reference into oracle/_ref/ (see oracle/build_ref.sh).
// No arithmetic lives here.
#pragma once
#include <cassert>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <numeric>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif
#define CV_32F 5
#define CV_64F 6
#define CV_64FC1 CV_64F

namespace cv
{
	struct DMatch
	{
		int queryIdx, trainIdx, imgIdx;
		float distance;
	};

	struct Size
	{
		int width, height;
		Size() : width(0), height(0) {}
		Size(int w, int h) : width(w), height(h) {}
	};

	// Row-major, contiguous, double-only matrix container (owning or viewing).
	class Mat
	{
	public:
		int rows, cols;
		unsigned char *data;

		Mat() : rows(0), cols(0), data(nullptr), type_(CV_64F) {}
		Mat(int r, int c, int t) : rows(r), cols(c), type_(t),
			storage_(std::make_shared<std::vector<double>>(static_cast<size_t>(r) * c, 0.0))
		{
			data = reinterpret_cast<unsigned char *>(storage_->data());
		}
		Mat(Size s, int t) : Mat(s.height, s.width, t) {}
		Mat(int r, int c, int t, void *external) : rows(r), cols(c),
			data(reinterpret_cast<unsigned char *>(external)), type_(t) {}

		int type() const { return type_; }
		Size size() const { return Size(cols, rows); }

		Mat row(int i) const
		{
			Mat view;
			view.rows = 1;
			view.cols = cols;
			view.type_ = type_;
			// non-owning view (cv::Mat::row is O(1) too; the parent outlives every row view on the compiled path)
			view.data = data + sizeof(double) * static_cast<size_t>(i) * cols;
			return view;
		}

		template <class T> T *ptr(int i = 0) { return reinterpret_cast<T *>(data) + static_cast<size_t>(i) * cols; }
		template <class T> const T *ptr(int i = 0) const { return reinterpret_cast<const T *>(data) + static_cast<size_t>(i) * cols; }
		template <class T> T &at(int i, int j) { return reinterpret_cast<T *>(data)[static_cast<size_t>(i) * cols + j]; }
		template <class T> const T &at(int i, int j) const { return reinterpret_cast<const T *>(data)[static_cast<size_t>(i) * cols + j]; }
		template <class T> T &at(int j) { return reinterpret_cast<T *>(data)[j]; }
		template <class T> const T &at(int j) const { return reinterpret_cast<const T *>(data)[j]; }

		// container operations math_utils.h's (unused) JAMA class names: (re)allocate, drop, copy out
		void create(int r, int c, int t) { *this = Mat(r, c, t); }
		void release() { *this = Mat(); }
		void convertTo(Mat &dst, int t) const
		{
			dst = Mat(rows, cols, t);
			if (rows * cols) std::memcpy(dst.data, data, sizeof(double) * static_cast<size_t>(rows) * cols);
		}
		const Mat &getMat() const { return *this; }

	private:
		int type_;
		std::shared_ptr<std::vector<double>> storage_;
	};
	typedef const Mat &InputArray;
}
