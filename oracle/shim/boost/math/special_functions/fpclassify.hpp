// empty stand-in: the reference only needs this header to exist (EmpiricalDistribution.cpp)
