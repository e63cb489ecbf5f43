// Host `Tucker` value class (vlasovtucker_b200/host/tucker.h, the reference's API) driven through
// the operations the reference's solver and tests use — construction with (precision, maxRank),
// operator+, scalar *, Hadamard *, Compress, Sum, Reconstructed — on tensors read from a file; the
// results go to a binary dump that tests/test_host_cpu.py compares with the oracle's Tucker algebra.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "header.h"

using namespace VlasovTucker;

static Tensor3d ReadTensor(std::ifstream& in, int n0, int n1, int n2)
{
    Tensor3d t(n0, n1, n2);
    in.read((char*)t.data(), sizeof(double) * n0 * n1 * n2);
    return t;
}
static void Write(std::ofstream& out, const Tucker& t)
{
    const auto r = t.Ranks();
    const double ranks[3] = {(double)r[0], (double)r[1], (double)r[2]};
    out.write((const char*)ranks, sizeof(ranks));
    const Tensor3d x = t.Reconstructed();
    out.write((const char*)x.data(), sizeof(double) * x.size());
    const double s = t.Sum();
    out.write((const char*)&s, sizeof(s));
}

int main(int argc, char** argv)
{
    if (argc == 2 && std::string(argv[1]) == "errors") {
        // error behaviour of the algebra (tucker.cpp:193, 262): std::invalid_argument with these texts
        Tensor3d x(3, 3, 3), y(3, 4, 3);
        x.setZero();
        y.setZero();
        const Tucker tx(x), ty(y);
        try {
            (void)(tx + ty);
        } catch (const std::invalid_argument& e) {
            std::cout << "sum: " << e.what() << "\n";
        }
        try {
            (void)(tx * ty);
        } catch (const std::invalid_argument& e) {
            std::cout << "mult: " << e.what() << "\n";
        }
        return 0;
    }
    if (argc < 8) {
        std::cerr << "usage: tucker_dump in.bin out.bin n0 n1 n2 eps maxRank\n";
        return 2;
    }
    const int n0 = atoi(argv[3]), n1 = atoi(argv[4]), n2 = atoi(argv[5]);
    const double eps = atof(argv[6]);
    const int rmax = atoi(argv[7]);
    std::ifstream in(argv[1], std::ios::binary);
    std::ofstream out(argv[2], std::ios::binary);
    const Tensor3d a = ReadTensor(in, n0, n1, n2), b = ReadTensor(in, n0, n1, n2);
    Tucker ta(a, eps, rmax), tb(b);           // truncated HOSVD / exact
    Write(out, ta);
    Tucker sum = ta + 0.7 * tb;               // ranks add
    Write(out, sum);
    sum.Compress(eps, rmax);
    Write(out, sum);
    Tucker had = ta * tb;                     // Hadamard: ranks multiply
    Write(out, had);
    had -= 0.25 * ta;
    had.Compress(eps, rmax);
    Write(out, had);
    return 0;
}
