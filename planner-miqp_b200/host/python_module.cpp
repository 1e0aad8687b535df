// python_module.cpp -- pybind11 module `miqp`: the Python surface of the reference's test module
// (python/bindings/python_module.cpp:13-20: CplexWrapper, SolutionProperties, OptimizationStatus, WarmstartType,
// ParallelMode, ConvexifiedMap) over the B200 host classes.  Built by planner-miqp_b200/build.py:build_pymodule().
//
// Differences that a caller sees: ConvexifiedMap takes the road polygon as an (n, 2) array and the reference line as an
// (n, 2) array of x, y (or a BARK trajectory array: columns 1 and 2 are x and y) instead of BARK geometry objects; its first
// argument (the BARK parameter server) is accepted and ignored.  BehaviorMiqpAgent is bound inside BARK's own module in the
// reference (python/bindings/python_planner_miqp.cpp) and needs BARK; it is not part of this module.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "b200_wrapper.hpp"
#include "convexified_map.hpp"

namespace py = pybind11;
using miqp::common::map::ConvexifiedMap;
using miqp::planner::MatrixXd;
using miqp::planner::OptimizationStatus;
using miqp::planner::SolutionProperties;
using miqp::planner::cplex::CplexWrapper;

namespace {

MatrixXd to_matrix(const py::array_t<double, py::array::c_style | py::array::forcecast> &a) {
  if (a.ndim() != 2 || a.shape(1) < 2) throw std::invalid_argument("expected an (n, 2) array of x, y");
  MatrixXd m((int)a.shape(0), 2);
  auto r = a.unchecked<2>();
  for (py::ssize_t k = 0; k < a.shape(0); ++k) { m((int)k, 0) = r(k, 0); m((int)k, 1) = r(k, 1); }
  return m;
}
py::array_t<double> to_array(const MatrixXd &m) {
  py::array_t<double> a({(py::ssize_t)m.rows(), (py::ssize_t)2});
  auto w = a.mutable_unchecked<2>();
  for (int k = 0; k < m.rows(); ++k) { w(k, 0) = m(k, 0); w(k, 1) = m(k, 1); }
  return a;
}
py::dict to_dict(const miqp::common::map::PolygonMap &pm) {
  py::dict d;
  for (const auto &kv : pm) d[py::int_(kv.first)] = to_array(kv.second);
  return d;
}

}  // namespace

PYBIND11_MODULE(miqp, m) {
  m.doc() = "B200-native MIQP planner backend: Python surface of planner-miqp's `miqp` test module";

  py::class_<ConvexifiedMap, std::shared_ptr<ConvexifiedMap>>(m, "ConvexifiedMap")
      .def(py::init([](py::object /*params*/, const py::array_t<double, py::array::c_style | py::array::forcecast> &poly, double buffer_radius,
                       double max_simplify_dist, double buffer_reference, double buffer_for_merging_tolerance) {
             return std::make_shared<ConvexifiedMap>(to_matrix(poly), buffer_radius, max_simplify_dist, buffer_reference, buffer_for_merging_tolerance);
           }),
           py::arg("params"), py::arg("map_polygon"), py::arg("buffer_radius"), py::arg("max_simplify_dist"), py::arg("buffer_reference"),
           py::arg("buffer_for_merging_tolerance"))
      .def("Convert", &ConvexifiedMap::Convert)
      .def("GetIntersectingConvexPolygons",
           [](const ConvexifiedMap &cm, const py::array_t<double, py::array::c_style | py::array::forcecast> &ref) {
             if (ref.ndim() != 2 || ref.shape(1) < 2) throw std::invalid_argument("expected an (n, 2) array of x, y or a BARK trajectory array");
             const int cx = ref.shape(1) >= 3 ? 1 : 0, cy = cx + 1;   // BARK StateDefinition: TIME, X, Y, THETA, VEL
             std::vector<miqp::planner::Point2> pts;
             auto r = ref.unchecked<2>();
             for (py::ssize_t k = 0; k < ref.shape(0); ++k) pts.push_back({r(k, cx), r(k, cy)});
             return to_dict(cm.GetIntersectingConvexPolygons(pts));
           })
      .def("HasValidPolygon", &ConvexifiedMap::HasValidPolygon)
      .def_property_readonly("map_nonconvex_polygon", [](const ConvexifiedMap &cm) { return to_array(cm.GetMapNonConvexPolygon()); }, "input map polygon.")
      .def_property_readonly("map_convex_polygons", [](const ConvexifiedMap &cm) { return to_dict(cm.GetMapConvexPolygons()); }, "decomposed convex polygons.");

  py::class_<CplexWrapper>(m, "CplexWrapper")
      // (modfile, precision): the reference's shorthand for a solver that reads an OPL .dat file (src/cplex_wrapper.hpp:119-120)
      .def(py::init([](const char *modfile, int precision) { return new CplexWrapper("cplexmodel/", modfile, CplexWrapper::DATFILE, precision); }))
      .def("setParameterDatFileAbsolute", &CplexWrapper::setParameterDatFileAbsolute)
      .def("callCplex", &CplexWrapper::callCplex, py::arg("timestamp") = 0.0, py::call_guard<py::gil_scoped_release>())
      .def("setDebugOutputFilePath", &CplexWrapper::setDebugOutputFilePath)
      .def("setDebugOutputFilePrefix", &CplexWrapper::setDebugOutputFilePrefix)
      .def("setDebugOutputPrint", &CplexWrapper::setDebugOutputPrint)
      .def("getDebugOutputParameterFilePath", &CplexWrapper::getDebugOutputParameterFilePath)
      .def("getSolutionProperties", &CplexWrapper::getSolutionProperties)
      // beyond the reference's binding: the file formats of the solver class
      .def("exportModel", &CplexWrapper::exportModel)
      .def("writeMIPStarts", &CplexWrapper::writeMIPStarts)
      .def("readMIPStarts", &CplexWrapper::readMIPStarts)
      .def("lastError", &CplexWrapper::lastError);

  py::enum_<OptimizationStatus>(m, "OptimizationStatus", py::arithmetic())
      .value("SUCCESS", OptimizationStatus::SUCCESS)
      .value("FAILED_NO_SOLUT", OptimizationStatus::FAILED_NO_SOLUT)
      .value("FAILED_SEG_FAULT", OptimizationStatus::FAILED_SEG_FAULT)
      .value("FAILED_TIMEOUT", OptimizationStatus::FAILED_TIMEOUT)
      .export_values();

  py::class_<SolutionProperties>(m, "SolutionProperties")
      .def(py::init())
      .def_readwrite("objective", &SolutionProperties::objective)
      .def_readwrite("status", &SolutionProperties::status)
      .def_readwrite("gap", &SolutionProperties::gap)
      .def_readwrite("time", &SolutionProperties::time);

  py::enum_<MiqpPlannerWarmstartType>(m, "WarmstartType")
      .value("NO_WARMSTART", NO_WARMSTART)
      .value("RECEDING_HORIZON_WARMSTART", RECEDING_HORIZON_WARMSTART)
      .value("LAST_SOLUTION_WARMSTART", LAST_SOLUTION_WARMSTART)
      .value("BOTH_WARMSTART_STRATEGIES", BOTH_WARMSTART_STRATEGIES)
      .export_values();

  py::enum_<MiqpPlannerParallelMode>(m, "ParallelMode")
      .value("AUTO", AUTO)
      .value("DETERMINISTIC", DETERMINISTIC)
      .value("OPPORTUNISTIC", OPPORTUNISTIC)
      .export_values();
}
