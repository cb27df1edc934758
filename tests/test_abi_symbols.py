"""The C-ABI libraries load without a GPU and export every symbol the headers declare."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, text)))


def test_cuda_library_exports_header(pkg):
    lib = pkg.cuda_api.load_library()
    names = declared("dymu_cuda.h", "dymu_")
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), n
    assert set(pkg.cuda_api.CUDA_SYMBOLS) == set(names)


def test_planner_library_exports_header(pkg):
    plib = pkg.planner_lib()
    assert plib.impl == "b200"
    names = declared("dymu_planner_c.h", "dymu_planner_")
    for n in names:
        assert hasattr(plib.lib, n), n
    assert set(pkg.planner_api.PLANNER_SYMBOLS) == set(names)


def test_reference_library_exports_same_api(ref_lib, pkg):
    assert ref_lib.impl == "reference"
    for n in pkg.planner_api.PLANNER_SYMBOLS:
        assert hasattr(ref_lib.lib, n), n


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the product path must fail loudly, not compute on the CPU."""
    import subprocess, sys
    code = ("import os,sys; os.environ['CUDA_VISIBLE_DEVICES']=''; sys.path.insert(0, %r); "
            "import dymu_b200; p=dymu_b200.load(); pl=p.DyMuPathPlanner(1.0,1.5,2.0,1); "
            "print('INIT', pl.initGlobalLayer(1.0,0.1,32,32))" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
    assert "INIT False" in out


def test_product_does_not_touch_oracle():
    """Nothing under the package may import, link or execute oracle/."""
    pkg_dir = os.path.join(ROOT, "planning-path_planning_b200")
    for base, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "import oracle" not in text and "dymu_oracle" not in text, f
                assert "libdymu_ref" not in text and "orc_" not in text, f


def test_every_reference_method_is_declared(tmp_path):
    """A caller that names any public method of the reference class (H.hpp:471-608) still compiles
    against the drop-in header: member-function pointers with the reference's exact signatures."""
    import subprocess
    src = tmp_path / "names_every_method.cpp"
    src.write_text(r'''
#include "DyMu.hpp"
using namespace PathPlanning_lib;
typedef DyMuPathPlanner P;
typedef base::Waypoint W;
void (P::*m01)(globalNode*) = &P::calculateSlope;
void (P::*m02)(globalNode*, int, int) = &P::calculateNominalCost;
void (P::*m03)(globalNode*) = &P::smoothCost;
globalNode* (P::*m04)(uint, uint) = &P::getGlobalNode;
bool (P::*m05)(W) = &P::setGoal;
bool (P::*m06)(W) = &P::computeTotalCostMap;
bool (P::*m07)() = &P::computeEntireTotalCostMap;
bool (P::*m08)(globalNode*) = &P::isSafeNode;
bool (P::*m09)(globalNode*) = &P::isFullyClosedNode;
void (P::*m10)() = &P::resetTotalCostMap;
void (P::*m11)() = &P::resetGlobalNarrowBand;
void (P::*m12)(globalNode*) = &P::propagateGlobalNode;
globalNode* (P::*m13)() = &P::minCostGlobalNode;
globalNode* (P::*m14)(base::Pose2D) = &P::getNearestGlobalNode;
globalNode* (P::*m15)(W) = &P::getNearestGlobalNode;
std::vector<W> (P::*m16)(W) = &P::getPath;
bool (P::*m17)(W) = &P::computeGlobalPath;
W (P::*m18)(W&, double) = &P::computeNextGlobalWaypoint;
void (P::*m19)(globalNode*, double&, double&) = &P::gradientNode;
double (P::*m20)(double, double, double, double, double, double) = &P::interpolate;
std::string (P::*m21)(W) = &P::getLocomotionMode;
std::vector<std::vector<double>> (P::*m22)() = &P::getTotalCostMatrix;
std::vector<std::vector<double>> (P::*m23)() = &P::getGlobalCostMatrix;
std::vector<std::vector<double>> (P::*m24)() = &P::getHazardDensityMatrix;
std::vector<std::vector<double>> (P::*m25)() = &P::getTrafficabilityMatrix;
double (P::*m26)(W) = &P::getTotalCost;
void (P::*m27)(globalNode*) = &P::createLocalMap;
localNode* (P::*m28)(base::Pose2D) = &P::getLocalNode;
localNode* (P::*m29)(W) = &P::getLocalNode;
void (P::*m30)(globalNode*) = &P::subdivideGlobalNode;
bool (P::*m31)(W, base::samples::frame::Frame, double, std::vector<W>&, base::Time&) = &P::computeLocalPlanning;
void (P::*m32)() = &P::expandRisk;
localNode* (P::*m33)() = &P::maxRiskNode;
void (P::*m34)(localNode*) = &P::propagateRisk;
void (P::*m35)(localNode*) = &P::setHorizonCost;
double (P::*m36)(localNode*) = &P::getTotalCost;
localNode* (P::*m37)(W, W) = &P::computeLocalPropagation;
void (P::*m38)(localNode*) = &P::propagateLocalNode;
localNode* (P::*m39)(double, double) = &P::minCostLocalNode;
localNode* (P::*m40)(localNode*) = &P::minCostLocalNode;
std::vector<W> (P::*m41)(localNode*, W, double) = &P::getLocalPath;
bool (P::*m42)(W&, double) = &P::computeLocalWaypointGDM;
W (P::*m43)(localNode*) = &P::computeLocalWaypointDijkstra;
void (P::*m44)(localNode*, double&, double&) = &P::gradientNode;
bool (P::*m45)(uint) = &P::evaluatePath;
bool (P::*m46)(localNode*, uint&, uint&) = &P::isBlockingObstacle;
int (P::*m47)(W, uint) = &P::repairPath;
std::vector<std::vector<double>> (P::*m48)(W) = &P::getRiskMatrix;
std::vector<std::vector<double>> (P::*m49)(W) = &P::getDeviationMatrix;
int (P::*m50)() = &P::getReconnectingIndex;
bool (P::*m51)(int, int, std::vector<double>) = &P::initCoRaMethod;
int (P::*m52)(base::samples::RigidBodyState) = &P::getTerrain;
bool (P::*m53)(int, std::vector<double>) = &P::fillTerrainInfo;
std::vector<double> (P::*m54)() = &P::updateCost;
std::vector<double> (P::*m55)() = &P::computeCostRatio;
bool (P::*m56)(double, double, uint, uint, std::vector<double>) = &P::initGlobalLayer;
bool (P::*m57)(std::vector<std::vector<double>>) = &P::setCostMap;
bool (P::*m58)(std::vector<double>, std::vector<double>, std::vector<std::string>, std::vector<std::vector<double>>,
               std::vector<std::vector<double>>) = &P::computeCostMap;
int main() { P p(1.0, 1.5, 2.0, SWEEPING); return p.getReconnectingIndex(); }
''')
    pkgdir = os.path.join(ROOT, "planning-path_planning_b200")
    exe = tmp_path / "names_every_method"
    cmd = ["/usr/bin/g++", "-std=c++14", "-I", os.path.join(pkgdir, "src"), "-I", os.path.join(pkgdir, "shim"),
           "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", pkgdir, "-ldymu_b200",
           "-ldymu_cuda", "-Wl,-rpath," + pkgdir]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
