// Stand-in for boost::math::digamma (Boost is not in /root/reference).  Defined in ref_em_driver.cpp.
#pragma once
namespace boost { namespace math { double digamma(double x); } }
