// tests/cpp/shim_demo.cpp -- drives the FILTER shim the way the reference's callbacks do (main.cpp:232-260):
// IMU samples arrive one by one, detection frames arrive at 25 Hz.  Reads "imu.txt"/"image.txt"-format files and prints
// one data/fusion.txt row per frame.  Built and run by tests/test_gpu_cpp_shim.py.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "../../include/fbus_filter.hpp"

int main(int argc, char** argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: shim_demo imu.txt image.txt n_init use_iir [threaded]\n"); return 2; }
    const size_t n_init = std::strtoul(argv[3], nullptr, 10);
    const bool iir = std::atoi(argv[4]) != 0;
    // threaded = 1: the reference's own arrangement -- IMU samples through the static InputIMUData callback, a filter thread
    // (StartFilterThread) that initialises gravity / gyro bias and is woken by SetDetectionResultUpdated
    const bool threaded = argc > 5 && std::atoi(argv[5]) != 0;
    fbus_config cfg;
    fbus_config_default(&cfg);
    FBUSB200::FILTER filter(cfg, 0, iir);
    std::ifstream fi(argv[1]), fd(argv[2]);
    std::vector<FBUSB200::IMUData> imu;
    for (std::string line; std::getline(fi, line);) {
        std::istringstream ss(line);
        FBUSB200::IMUData d;
        if (ss >> d.timeStamp >> d.accel[0] >> d.accel[1] >> d.accel[2] >> d.gyro[0] >> d.gyro[1] >> d.gyro[2]) imu.push_back(d);
    }
    size_t next = 0;
    auto push = [&](const FBUSB200::IMUData& d) {
        if (threaded) FBUSB200::FILTER::InputIMUData(d, &filter);
        else filter.SetImuData(d);
    };
    for (; next < n_init && next < imu.size(); ++next) push(imu[next]);
    if (threaded) {
        filter.StartFilterThread(50);  // the reference waits 1000 ms for IMU data (filter.cpp:193); the data is already there
        filter.WaitIdle();
    } else {
        filter.InitializeGravityAndBias();  // the reference does this 1 s after start-up (filter.cpp:193-196)
    }
    size_t n_frames = 0, n_viewer_ok = 0;
    FBUSB200::DetectionResultList frame;
    auto flush = [&]() {
        if (frame.empty()) return;
        // the IMU callback keeps running while the vision thread works: a few samples later than the frame are
        // already buffered when the filter wakes up (they stay buffered, filter.cpp:501-503)
        while (next < imu.size() && imu[next].timeStamp <= frame[0].timeStamp) push(imu[next++]);
        for (int extra = 0; extra < 3 && next < imu.size(); ++extra) push(imu[next++]);
        filter.SetDetectionResult(frame);
        filter.SetDetectionResultUpdated();
        if (threaded) filter.WaitIdle();
        const auto r = filter.GetFusionRow();
        {   // what the viewer thread asks for (visualizer.cpp:49) agrees with the single getters
            FBUSB200::Matrix4d cam, vis;
            std::vector<FBUSB200::Matrix4d> st, dy, st2;
            filter.GetVisualizeInfo(cam, vis, st, dy);
            filter.GetStaticMarkerPose(st2);
            ++n_frames;
            if (cam == filter.GetCameraPose() && vis == filter.GetVisualPose() && st == dy && st == st2 &&
                st.size() == (cam[15] != 0 ? (size_t)12 : (size_t)0) && cam[3] == r[1] && cam[7] == r[2] && cam[11] == r[3])
                ++n_viewer_ok;
        }
        for (size_t i = 0; i < r.size(); ++i) std::printf("%.17g%c", r[i], i + 1 == r.size() ? '\n' : ' ');
        frame.clear();
    };
    for (std::string line; std::getline(fd, line);) {
        std::istringstream ss(line);
        FBUSB200::DetectionResult d;
        double id;
        if (!(ss >> d.timeStamp >> id >> d.positionAtCL[0] >> d.positionAtCL[1] >> d.positionAtCL[2] >> d.quaternionM2CL[0] >>
              d.quaternionM2CL[1] >> d.quaternionM2CL[2] >> d.quaternionM2CL[3]))
            continue;
        d.markerID = (int)id;
        if (!frame.empty() && frame[0].timeStamp != d.timeStamp) flush();
        frame.push_back(d);
    }
    flush();
    if (threaded) filter.StopFilterThread();
    std::fprintf(stderr, "viewer-info %zu/%zu\n", n_viewer_ok, n_frames);
    return n_viewer_ok == n_frames ? 0 : 3;
}
