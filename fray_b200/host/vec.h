// vec.h -- FP64 vector / 3x3 matrix / transform and FP32 colour types of the host scene layer.
//
// Semantics follow the reference's math layer so that the flattened scene carries the same numbers:
//   Vector      /root/reference/src/vector.h:30-213   (double x,y,z)
//   Matrix      /root/reference/src/matrix.h:28-70, src/matrix.cpp:29-108 (row vector * matrix)
//   Transform   /root/reference/src/matrix.h:72-98,  src/matrix.cpp:110-161
//   Color       /root/reference/src/color.h:37-170    (float r,g,b)
#pragma once
#include <cmath>
#include <algorithm>

namespace fray {

constexpr double kPi = 3.141592653589793238; // PI, src/constants.h:31
constexpr double kInf = 1e99;                // INF, src/constants.h:32

inline double radians(double deg) { return deg / 180.0 * kPi; } // toRadians, src/util.h:36

struct Vec3 {
	double x = 0, y = 0, z = 0;
	Vec3() {}
	Vec3(double x, double y, double z): x(x), y(y), z(z) {}
	double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
	double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
	double lengthSqr() const { return x * x + y * y + z * z; }
	double length() const { return std::sqrt(lengthSqr()); }
	// normalise by multiplying with the reciprocal length, as src/vector.h:84-88 does
	void normalize() { double k = 1.0 / length(); x *= k; y *= k; z *= k; }
};
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator-(const Vec3& a) { return Vec3(-a.x, -a.y, -a.z); }
inline Vec3 operator*(const Vec3& a, double k) { return Vec3(a.x * k, a.y * k, a.z * k); }
inline Vec3 operator*(double k, const Vec3& a) { return Vec3(a.x * k, a.y * k, a.z * k); }
inline double dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(const Vec3& a, const Vec3& b)
{
	return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline Vec3 normalized(Vec3 v) { v.normalize(); return v; }

struct Mat3 {
	double m[3][3];
	Mat3() { *this = diag(1); }
	static Mat3 diag(double d)
	{
		Mat3 r(0);
		r.m[0][0] = r.m[1][1] = r.m[2][2] = d;
		return r;
	}
private:
	explicit Mat3(int) { for (auto& row: m) for (double& e: row) e = 0; }
	friend Mat3 operator*(const Mat3&, const Mat3&);
	friend Mat3 inverse(const Mat3&);
};

// row vector times matrix, src/matrix.h:53-60
inline Vec3 operator*(const Vec3& v, const Mat3& a)
{
	return Vec3(v.x * a.m[0][0] + v.y * a.m[1][0] + v.z * a.m[2][0],
	            v.x * a.m[0][1] + v.y * a.m[1][1] + v.z * a.m[2][1],
	            v.x * a.m[0][2] + v.y * a.m[1][2] + v.z * a.m[2][2]);
}

inline Mat3 operator*(const Mat3& a, const Mat3& b) // src/matrix.cpp:65-73 (k innermost, accumulating from 0)
{
	Mat3 c(0);
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++)
			for (int k = 0; k < 3; k++) c.m[i][j] += a.m[i][k] * b.m[k][j];
	return c;
}

inline double determinant(const Mat3& a) // term order of src/matrix.cpp:75-83
{
	return a.m[0][0] * a.m[1][1] * a.m[2][2] - a.m[0][0] * a.m[1][2] * a.m[2][1]
	     - a.m[0][1] * a.m[1][0] * a.m[2][2] + a.m[0][1] * a.m[1][2] * a.m[2][0]
	     + a.m[0][2] * a.m[1][0] * a.m[2][1] - a.m[0][2] * a.m[1][1] * a.m[2][0];
}

inline Mat3 inverse(const Mat3& a) // adjugate / determinant, src/matrix.cpp:85-108
{
	double D = determinant(a);
	if (std::fabs(D) < 1e-12) return a; // the reference returns the input when singular
	double rD = 1.0 / D;
	Mat3 r(0);
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) {
			// cofactor of element (j, i)
			int r0 = (j == 0) ? 1 : 0, r1 = (j == 2) ? 1 : 2;
			int c0 = (i == 0) ? 1 : 0, c1 = (i == 2) ? 1 : 2;
			double t = a.m[r0][c0] * a.m[r1][c1] - a.m[r1][c0] * a.m[r0][c1];
			if ((i + j) % 2) t = -t;
			r.m[i][j] = rD * t;
		}
	return r;
}

inline Mat3 rotX(double a) // src/matrix.cpp:29-39
{
	double S = std::sin(a), C = std::cos(a);
	Mat3 r;
	r.m[1][1] = C; r.m[2][1] = S; r.m[1][2] = -S; r.m[2][2] = C;
	return r;
}
inline Mat3 rotY(double a) // src/matrix.cpp:41-51
{
	double S = std::sin(a), C = std::cos(a);
	Mat3 r;
	r.m[0][0] = C; r.m[2][0] = -S; r.m[0][2] = S; r.m[2][2] = C;
	return r;
}
inline Mat3 rotZ(double a) // src/matrix.cpp:53-63
{
	double S = std::sin(a), C = std::cos(a);
	Mat3 r;
	r.m[0][0] = C; r.m[1][0] = S; r.m[0][1] = -S; r.m[1][1] = C;
	return r;
}

struct Transform {
	Vec3 offset;
	Mat3 m, inv;
	void scale(double x, double y, double z) // src/matrix.cpp:117-127
	{
		Mat3 s = Mat3::diag(0);
		s.m[0][0] = x; s.m[1][1] = y; s.m[2][2] = z;
		m = m * s;
		inv = inverse(m);
	}
	void rotate(double yaw, double pitch, double roll) // src/matrix.cpp:129-135
	{
		m = m * rotZ(radians(roll)) * rotX(radians(pitch)) * rotY(radians(yaw));
		inv = inverse(m);
	}
	void translate(const Vec3& t) { offset = offset + t; } // src/matrix.cpp:137-140
	Vec3 point(const Vec3& p) const { return p * m + offset; }         // transformPoint
	Vec3 unpoint(const Vec3& p) const { return (p - offset) * inv; }    // untransformPoint
	Vec3 dir(const Vec3& d) const { return normalized(d * m); }         // transformDir
	Vec3 undir(const Vec3& d) const { return normalized(d * inv); }     // untransformDir
};

struct Color {
	float r = 0, g = 0, b = 0;
	Color() {}
	Color(float r, float g, float b): r(r), g(g), b(b) {}
	float intensity() const { return (r + g + b) / 3; } // src/color.h:81-84
};

} // namespace fray
