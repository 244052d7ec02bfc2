"""Builds the reference's OWN wrapper sources and offline drivers, unchanged, against this repository's backend.

    python compat/build_ref.py            (also called by __graft_entry__.build() when /root/reference is present)

Inputs, read where they lie (never copied into the repository):
    $REF/gtsam/{gtsam_graph,imu_base,imu_vn100,gt_parameter,color}.cpp    -> compat/_ref/libgraphslam_gt.so
    $REF/gtsam/{test_vro_imu_graph,test_ba_imu_graph}.cpp                 -> compat/_ref/<driver>           (their own main())
    $REF/g2o/{g2o_graph,g2o_parameter,color}.cpp                          -> compat/_ref/libgraphslam_g2o.so
    tests/cpp/*.cpp (this repository's small test programs, written against the reference's headers) -> compat/_ref/<name>
compiled with -Icompat (Eigen / ROS / OpenCV / Qt / Boost / vro stand-ins, compat/README.md), the gtsam:: and g2o:: facades
(graph_slam_b200/host/gtsam_lite.h, g2o_lite.h) and linked with graph_slam_b200/libfg_b200.so.  compat/_ref/ is git-ignored
and travels to the GPU box with the snapshot, where /root/reference does not exist: the GPU tests run these binaries."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('GRAPH_SLAM_REF', '/root/reference')
OUT = os.path.join(ROOT, 'compat', '_ref')
LIBDIR = os.path.join(ROOT, 'graph_slam_b200')
GT_SRC = ['gtsam_graph.cpp', 'imu_base.cpp', 'imu_vn100.cpp', 'gt_parameter.cpp', 'color.cpp']
GT_DRIVERS = ['test_vro_imu_graph', 'test_ba_imu_graph']
G2O_SRC = ['g2o_graph.cpp', 'g2o_parameter.cpp', 'color.cpp']
TEST_PROGRAMS = {'vio_driver': 'gt', 'ba_driver': 'gt', 'format_io': 'gt', 'plane_driver': 'gt', 'plane_check': 'gt', 'g2o_driver': 'g2o'}


def available():
    return os.path.isfile(os.path.join(REF, 'gtsam', 'gtsam_graph.cpp'))


def flags(sub):
    inc = [os.path.join(ROOT, 'compat'), os.path.join(ROOT, 'compat', 'vro'), os.path.join(ROOT, 'compat', 'qt'), os.path.join(ROOT, 'include'),
           os.path.join(REF, sub)]
    # -fpermissive -w: the reference's own build flags (CMakeLists.txt:20)
    return ['-std=c++17', '-O1', '-fpermissive', '-w', '-fPIC'] + ['-I' + i for i in inc]


def newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def common_deps():
    deps = [os.path.abspath(__file__), os.path.join(LIBDIR, 'host', 'gtsam_lite.h'), os.path.join(LIBDIR, 'host', 'g2o_lite.h'),
            os.path.join(ROOT, 'include', 'fg_abi.h')]
    for d, _, files in os.walk(os.path.join(ROOT, 'compat')):
        if '_ref' in d:
            continue
        deps += [os.path.join(d, f) for f in files]
    return deps


def run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('compat/build_ref.py: %s\n%s' % (' '.join(cmd[:6]) + ' ...', res.stderr[-4000:]))


def build(force=False, verbose=False):
    """Returns the list of artefacts built (or already up to date); [] when the reference is not on this machine."""
    if not available():
        return []
    os.makedirs(OUT, exist_ok=True)
    link = ['-L' + LIBDIR, '-lfg_b200', '-Wl,-rpath,$ORIGIN/../../graph_slam_b200', '-Wl,-rpath,$ORIGIN']
    deps = common_deps()
    done = []
    libs = {}
    for sub, srcs, name in (('gtsam', GT_SRC, 'libgraphslam_gt.so'), ('g2o', G2O_SRC, 'libgraphslam_g2o.so')):
        src = [os.path.join(REF, sub, s) for s in srcs]
        if not all(os.path.isfile(s) for s in src):
            continue
        lib = os.path.join(OUT, name)
        if force or newer(lib, deps + src):
            if verbose:
                print('building', lib, flush=True)
            run(['g++'] + flags(sub) + ['-shared'] + src + link + ['-o', lib])
        libs[sub] = lib
        done.append(lib)
    for d in GT_DRIVERS:
        src = os.path.join(REF, 'gtsam', d + '.cpp')
        exe = os.path.join(OUT, d)
        if 'gtsam' in libs and os.path.isfile(src) and (force or newer(exe, deps + [src, libs['gtsam']])):
            run(['g++'] + flags('gtsam') + [src, '-L' + OUT, '-lgraphslam_gt'] + link + ['-o', exe])
        if os.path.exists(exe):
            done.append(exe)
    for prog, kind in TEST_PROGRAMS.items():
        src = os.path.join(ROOT, 'tests', 'cpp', prog + '.cpp')
        sub, libname = ('gtsam', 'graphslam_gt') if kind == 'gt' else ('g2o', 'graphslam_g2o')
        exe = os.path.join(OUT, prog)
        if sub in libs and os.path.isfile(src) and (force or newer(exe, deps + [src, libs[sub]])):
            run(['g++'] + flags(sub) + [src, '-L' + OUT, '-l' + libname] + link + ['-o', exe])
        if os.path.exists(exe):
            done.append(exe)
    return done


if __name__ == '__main__':
    if not available():
        print('reference not found at %s: nothing to build (prebuilt files in compat/_ref/ are used as they are)' % REF)
        sys.exit(0)
    for a in build(force='--force' in sys.argv, verbose=True):
        print(a)
