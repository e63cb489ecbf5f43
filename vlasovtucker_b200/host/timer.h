// Wall-clock section timer printing to stderr (reference: src/timer.h:9-48).
#pragma once
#include <chrono>
#include <iostream>
#include <string>

namespace VlasovTucker {
class Timer {
    using clock = std::chrono::steady_clock;

public:
    Timer() : t0_(clock::now()), last_(t0_) {}
    ~Timer() { std::cerr << "Total execution time: " << seconds(t0_) << "s\n"; }
    void StartSection() { last_ = clock::now(); }
    void PrintSectionTime(const std::string& name = "")
    {
        const double s = seconds(last_);
        last_ = clock::now();
        if (!name.empty()) std::cerr << name << ": ";
        std::cerr << s << "s\n";
    }

private:
    static double seconds(clock::time_point since)
    {
        return std::chrono::duration_cast<std::chrono::microseconds>(clock::now() - since).count() / 1.0e6;
    }
    clock::time_point t0_, last_;
};
}  // namespace VlasovTucker
