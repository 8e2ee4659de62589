/* Complex.h -- double-precision complex number with the interface the
 * reference's callers use (reference Complex.h:19-342: class Complex with
 * public re/im, arithmetic, conj/abs2/expi/realProduct helpers and "(re,im)"
 * stream format).  Written for fftwpp-b200; layout-compatible with
 * double[2], std::complex<double> and CUDA double2 (16 bytes, re first), which
 * is what the device kernels read.
 */
#ifndef __Complex_h__
#define __Complex_h__ 1

#include <cmath>
#include <iostream>

class Complex {
public:
  double re,im;

  Complex() {}
  Complex(double r, double i=0.0) : re(r), im(i) {}

  double real() const {return re;}
  double imag() const {return im;}

  Complex& operator+=(const Complex& z) {re += z.re; im += z.im; return *this;}
  Complex& operator-=(const Complex& z) {re -= z.re; im -= z.im; return *this;}
  Complex& operator+=(double x) {re += x; return *this;}
  Complex& operator-=(double x) {re -= x; return *this;}
  Complex& operator*=(double x) {re *= x; im *= x; return *this;}
  Complex& operator/=(double x) {re /= x; im /= x; return *this;}
  Complex& operator*=(const Complex& z) {
    const double a=re*z.re-im*z.im;
    im=re*z.im+im*z.re;
    re=a;
    return *this;
  }
  Complex& operator/=(const Complex& z) {
    const double s=1.0/(z.re*z.re+z.im*z.im);
    const double a=(re*z.re+im*z.im)*s;
    im=(im*z.re-re*z.im)*s;
    re=a;
    return *this;
  }
};

inline bool operator==(const Complex& a, const Complex& b) {return a.re == b.re && a.im == b.im;}
inline bool operator==(const Complex& a, double b) {return a.re == b && a.im == 0.0;}
inline bool operator!=(const Complex& a, const Complex& b) {return !(a == b);}
inline bool operator!=(const Complex& a, double b) {return !(a == b);}

inline Complex operator-(const Complex& a) {return Complex(-a.re,-a.im);}
inline Complex conj(const Complex& a) {return Complex(a.re,-a.im);}

inline Complex operator+(const Complex& a, const Complex& b) {return Complex(a.re+b.re,a.im+b.im);}
inline Complex operator+(const Complex& a, double b) {return Complex(a.re+b,a.im);}
inline Complex operator+(double a, const Complex& b) {return Complex(a+b.re,b.im);}
inline Complex operator-(const Complex& a, const Complex& b) {return Complex(a.re-b.re,a.im-b.im);}
inline Complex operator-(const Complex& a, double b) {return Complex(a.re-b,a.im);}
inline Complex operator-(double a, const Complex& b) {return Complex(a-b.re,-b.im);}
inline Complex operator*(const Complex& a, const Complex& b)
{
  return Complex(a.re*b.re-a.im*b.im,a.re*b.im+a.im*b.re);
}
inline Complex operator*(const Complex& a, double b) {return Complex(a.re*b,a.im*b);}
inline Complex operator*(double a, const Complex& b) {return Complex(a*b.re,a*b.im);}
// a*conj(b)
inline Complex multconj(const Complex& a, const Complex& b)
{
  return Complex(a.re*b.re+a.im*b.im,a.im*b.re-a.re*b.im);
}
inline Complex operator/(const Complex& a, const Complex& b) {Complex q(a); q /= b; return q;}
inline Complex operator/(const Complex& a, double b) {return Complex(a.re/b,a.im/b);}
inline Complex operator/(double a, const Complex& b) {Complex q(a); q /= b; return q;}

inline double real(const Complex& a) {return a.re;}
inline double imag(const Complex& a) {return a.im;}
inline double abs2(const Complex& a) {return a.re*a.re+a.im*a.im;}
inline double abs(const Complex& a) {return std::hypot(a.re,a.im);}
inline double arg(const Complex& a) {return std::atan2(a.im,a.re);}
inline Complex polar(double r, double t) {return Complex(r*std::cos(t),r*std::sin(t));}

inline Complex sqrt(const Complex& a)
{
  if(a.re == 0.0 && a.im == 0.0) return Complex(0.0,0.0);
  const double r=abs(a);
  const double s=std::sqrt(0.5*(r+std::fabs(a.re)));
  const double t=0.5*a.im/s;
  if(a.re > 0.0) return Complex(s,t);
  return a.im >= 0.0 ? Complex(std::fabs(t),s) : Complex(std::fabs(t),-s);
}

inline double realProduct(double a, double b) {return a*b;}
// Re(conj(a)*b)
inline double realProduct(const Complex& a, const Complex& b) {return a.re*b.re+a.im*b.im;}

inline Complex expi(double phase) {return Complex(std::cos(phase),std::sin(phase));}
inline Complex exp(const Complex& a) {return std::exp(a.re)*expi(a.im);}
inline Complex pow(const Complex& a, double u)
{
  if(a == 0.0) return u == 0.0 ? 1.0 : 0.0;
  return polar(std::pow(abs2(a),0.5*u),u*arg(a));
}
inline Complex pow(const Complex& a, const Complex& w)
{
  if(a == 0.0) return w == 0.0 ? 1.0 : 0.0;
  const double lr=0.5*std::log(abs2(a)), th=arg(a);
  return polar(std::exp(lr*w.re-th*w.im),lr*w.im+th*w.re);
}

// "(re,im)" on output; "(re,im)", "(re)" or "re" on input
inline std::ostream& operator<<(std::ostream& s, const Complex& a)
{
  return s << "(" << a.re << "," << a.im << ")";
}
inline std::istream& operator>>(std::istream& s, Complex& a)
{
  char c=0;
  a.im=0.0;
  s >> std::ws;
  if(s.peek() != '(') return s >> a.re;
  s >> c >> a.re >> c;
  if(c == ',') s >> a.im >> c;
  return s;
}

#endif
