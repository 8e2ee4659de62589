/* seconds.h -- wall-clock timers with the reference's names (reference
 * seconds.h:43-100: utils::stopWatch, utils::cpuTimer, utils::cpuTime).
 * DEVIATION: the reference's cpuTimer reports min(wall, process CPU time);
 * here the arithmetic runs on the GPU while the host thread waits, so CPU
 * time says nothing about the convolution -- cpuTimer reports wall time.
 */
#ifndef __seconds_h__
#define __seconds_h__ 1

#include <chrono>
#include <cstdio>
#include <ctime>

namespace utils {

inline double cpuTime()
{
  timespec t;
  clock_gettime(CLOCK_PROCESS_CPUTIME_ID,&t);
  return 1.0e9*t.tv_sec+t.tv_nsec;
}

class stopWatch {
  typedef std::chrono::steady_clock clock;
  clock::time_point t0;
public:
  stopWatch() {reset();}
  void reset() {t0=clock::now();}
  double nanoseconds(bool restart=false) {
    clock::time_point t1=clock::now();
    double ns=std::chrono::duration<double,std::nano>(t1-t0).count();
    if(restart) t0=t1;
    return ns;
  }
  double seconds(bool restart=false) {return 1.0e-9*nanoseconds(restart);}
};

class cpuTimer : public stopWatch {};

}

inline int renameOverwrite(const char *oldpath, const char *newpath)
{
  return rename(oldpath,newpath);
}

#endif
