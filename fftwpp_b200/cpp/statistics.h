/* statistics.h -- running sample statistics with the reference's interface
 * (reference statistics.h:19-127: utils::statistics with add/mean/min/max/
 * stdev/stderror/median/output).  Samples are kept, so the median is exact.
 */
#ifndef __statistics_h__
#define __statistics_h__ 1

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <vector>

namespace utils {

class statistics {
  std::vector<double> t;
  bool wantMedian;
  double moment(int side, double about) const {
    double v=0.0;
    for(size_t i=0; i < t.size(); ++i) {
      const double d=t[i]-about;
      if(side == 0 || (side < 0 ? d < 0.0 : d >= 0.0)) v += d*d;
    }
    return v;
  }
  double dev(double var, double f) const {
    return t.size() <= f ? DBL_MAX : std::sqrt(var*f/(t.size()-f));
  }
public:
  statistics(bool computeMedian=false) : wantMedian(computeMedian) {}
  void clear() {t.clear();}
  void add(double x) {t.push_back(x);}
  double count() {return (double) t.size();}
  double sum() {double s=0.0; for(size_t i=0; i < t.size(); ++i) s += t[i]; return s;}
  double mean() {return t.empty() ? 0.0 : sum()/t.size();}
  double min() {return t.empty() ? DBL_MAX : *std::min_element(t.begin(),t.end());}
  double max() {return t.empty() ? -DBL_MAX : *std::max_element(t.begin(),t.end());}
  double stdev() {return dev(moment(0,mean()),1.0);}
  double stdevL() {return dev(moment(-1,mean()),2.0);}
  double stdevH() {return dev(moment(1,mean()),2.0);}
  double stderror() {return stdev()/std::sqrt((double) t.size());}
  double median() {
    if(!wantMedian) {
      std::cerr << "Constructor requires median=true" << std::endl;
      exit(-1);
    }
    std::vector<double> s(t);
    std::sort(s.begin(),s.end());
    const size_t h=s.size()/2;
    return s.empty() ? 0.0 : (2*h == s.size() ? 0.5*(s[h-1]+s[h]) : s[h]);
  }
  void output(const char *text, size_t m) {
    std::cout << text << ": \n" << m << "\t" << mean() << "\t" << stdevL()
              << "\t" << stdevH() << std::endl;
  }
};

}

#endif
