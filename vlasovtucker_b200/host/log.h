// Console / log.txt sink used by the solvers (reference: src/log.h:10-79).
#pragma once
#include <fstream>
#include <iostream>
#include <string>

namespace VlasovTucker {
enum class LogLevel { None, Console, TextFile, AllText };

class Log {
public:
    Log() = default;
    explicit Log(LogLevel level, std::string prefix = "") : level_(level), prefix_(std::move(prefix)) { reopen(); }
    Log(const Log& o) : level_(o.level_), prefix_(o.prefix_) { reopen(); }
    Log& operator=(const Log& o)
    {
        if (this != &o) {
            file_.close();
            level_ = o.level_;
            prefix_ = o.prefix_;
            reopen();
        }
        return *this;
    }
    template <typename T>
    Log& operator<<(const T& x)
    {
        if (level_ == LogLevel::Console || level_ == LogLevel::AllText) std::cout << x;
        if (toFile() && file_.is_open()) file_ << x << std::flush;
        return *this;
    }

private:
    bool toFile() const { return level_ == LogLevel::TextFile || level_ == LogLevel::AllText; }
    void reopen()
    {
        if (toFile()) file_.open(prefix_ + "log.txt");
    }
    LogLevel level_ = LogLevel::None;
    std::string prefix_;
    std::ofstream file_;
};

inline std::string Indent(int level) { return std::string(4 * (size_t)level, ' '); }
}  // namespace VlasovTucker
