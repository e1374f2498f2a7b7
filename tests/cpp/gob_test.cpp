// Wire-format check of include/paillier_b200.hpp against paillier_b200/gobwire.py (driven by tests/test_cpp_host_mirror.py).
// stdin: lines "enc <C hex> <level> <method>" -> prints the stream as hex; "dec <stream hex>" -> prints "<C hex> <level> <method>"
// or "error <message>".
#include <iostream>
#include <string>

#include "paillier_b200.hpp"

using namespace paillier;

static std::string hex_of(const std::vector<uint8_t>& b) {
    static const char* d = "0123456789abcdef";
    std::string s;
    for (uint8_t x : b) { s += d[x >> 4]; s += d[x & 15]; }
    return s;
}

static std::vector<uint8_t> raw_from_hex(const std::string& h) {
    std::vector<uint8_t> out;
    for (size_t i = 0; i + 1 < h.size(); i += 2) out.push_back((uint8_t)std::stoul(h.substr(i, 2), nullptr, 16));
    return out;
}

int main() {
    std::string kind;
    while (std::cin >> kind) {
        if (kind == "enc") {
            std::string c; int level, method;
            std::cin >> c >> level >> method;
            Ciphertext ct; ct.C = from_hex(c); ct.Level = level; ct.EncMethod = method;
            std::cout << hex_of(ct.Bytes()) << "\n";
        } else if (kind == "dec") {
            std::string h; std::cin >> h;
            if (h == "-") h.clear();
            try {
                Ciphertext ct = NewCiphertextFromBytes(raw_from_hex(h));
                std::cout << to_hex(ct.C) << " " << ct.Level << " " << ct.EncMethod << "\n";
            } catch (const Error& e) {
                std::cout << "error " << e.what() << "\n";
            }
        } else {
            return 2;
        }
    }
    return 0;
}
