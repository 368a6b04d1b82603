/* proto.i -- SWIG interface for hosts that have swig (this image does not; the pybind11 twin in
 * proto_pybind.cpp is what build() compiles).  Same module and template names as the reference's
 * src/cpp/proto.i:1-27; only the included headers differ: the engine-backed facade. */
%module proto

%{
#include <partitionsolvers/DP.hpp>
#include "python_dpsolver.hpp"
#include <partitionsolvers/LTSS.hpp>
#include "python_ltsssolver.hpp"
%}

%include "std_vector.i"
%include "std_pair.i"
%include "exception.i"

%exception {
  try { $action }
  catch (const distributionException& e) { SWIG_exception(SWIG_ValueError, e.what()); }
  catch (const std::exception& e) { SWIG_exception(SWIG_RuntimeError, e.what()); }
}

namespace std {
%template(IArray) vector<int>;
%template(FArray) vector<float>;
%template(IArrayArray) vector<vector<int> >;
%template(IArrayFPair) pair<vector<int>, float>;
%template(IArrayArrayFPair) pair<vector<vector<int> >, float>;
%template(SWCont) vector<pair<vector<vector<int> >, float> >;
%template(OLSPair) pair<vector<vector<int> >, int >;
%template(ScoreCont) pair<vector<pair<vector<vector<int> >,float> >, int>;
}

%include "python_dpsolver.hpp"
%include "python_ltsssolver.hpp"
