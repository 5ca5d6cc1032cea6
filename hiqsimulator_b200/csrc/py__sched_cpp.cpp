// pybind11 module `_sched_cpp`: same two classes, constructor signatures and methods as the
// reference binding (reference: _sched_cpp.cpp:28-39), backed by csrc/sched.cpp.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "sched.hpp"

namespace py = pybind11;
using namespace hiq::sched;

PYBIND11_MODULE(_sched_cpp, m)
{
     m.doc() = "B200 engine host scheduler: drop-in for HiQsimulator's _sched_cpp";
     py::class_<SwapScheduler>(m, "SwapScheduler")
         .def(py::init<const std::vector<std::vector<Id>>&, const std::vector<std::vector<Id>>&, std::vector<bool>, int, int, bool>())
         .def("ScheduleSwap", &SwapScheduler::ScheduleSwap, py::call_guard<py::gil_scoped_release>());
     py::class_<ClusterScheduler>(m, "ClusterScheduler")
         .def(py::init<const std::vector<std::vector<Id>>&, const std::vector<std::vector<Id>>&, std::vector<bool>,
                       const std::vector<Id>&, const std::vector<Id>&, int>())
         .def("ScheduleCluster", &ClusterScheduler::ScheduleCluster, py::call_guard<py::gil_scoped_release>())
         .def("candidates", &ClusterScheduler::candidates)
         .def("evaluated", &ClusterScheduler::evaluated);
     py::class_<GreedyPlanner>(m, "GreedyPlanner",
                               "stage / cluster loop of the reference's GreedyScheduler as one object: next() -> (kind, ids) with kind "
                               "0 done, 1 perm, 2 controlled-Z role swap (gate, control position), 3 cluster (gate indices), 4 swap pairs")
         .def(py::init<std::vector<std::vector<Id>>, std::vector<std::vector<Id>>, std::vector<bool>, std::vector<Id>, std::vector<Id>, int,
                       int, bool>())
         .def("next",
              [](GreedyPlanner& p) {
                   GreedyPlanner::Step s;
                   {
                        py::gil_scoped_release release;
                        s = p.next();
                   }
                   return py::make_tuple(s.kind, s.data);
              })
         .def("cluster_seconds", &GreedyPlanner::cluster_seconds)
         .def("swap_seconds", &GreedyPlanner::swap_seconds);
     m.def("set_mode", &ClusterScheduler::set_mode, "0 = bounded search (default), 1 = replay the reference's enumeration and score every candidate");
     m.def("set_threads", &ClusterScheduler::set_threads, "host threads used to score candidate clusters (0 = auto)");
}
