// Compile-and-link check of include/variational_mt_gpu.hpp: the call sites of slow_flow.cpp:673,875-888,1018-1023
// written against a minimal ParameterList with the reference's lookup signatures (utils/parameter_list.h:20-143).
#include <map>
#include <sstream>
#include <string>

#include "variational_mt_gpu.hpp"

class ParameterList {
public:
    void insert(std::string k, std::string v, bool overwrite = false) { if (overwrite || !m_.count(k)) m_[k] = v; }
    bool exists(std::string k) { return m_.count(k) != 0; }
    std::string parameter(const char *k) { return m_[k]; }
    template <typename T> T parameter(std::string k, std::string def) {
        std::istringstream s(exists(k) ? m_[k] : def);
        T v = T();
        s >> v;
        return v;
    }
private:
    std::map<std::string, std::string> m_;
};
template <> inline bool ParameterList::parameter<bool>(std::string k, std::string def) { return (exists(k) ? m_[k] : def) != "0"; }

int main(int argc, char **) {
    ParameterList thread_params;
    thread_params.insert("slow_flow_S", "3", true);
    if (argc > 1000) { // never executed by the test: this only has to compile and link
        color_image_t *seq[5] = {0, 0, 0, 0, 0};
        image_t *wx = 0, *wy = 0;
        normalize(seq, 5, thread_params);                       // slow_flow.cpp:673
        Variational_MT minimzer_f;                               // slow_flow.cpp:875-888
        minimzer_f.setChannelWeights(0);
        sf_point2f r = minimzer_f.variational(wx, wy, seq, thread_params);
        image_t *occlusions = minimzer_f.getOcclusions();
        Variational_MT minimzer_b;                               // slow_flow.cpp:1018-1023
        minimzer_b.one_direction = true;
        minimzer_b.variational(wx, wy, seq, thread_params);
        variational_params_t flow_params;                        // adaptiveFR.cpp:293-302, 574
        variational_params_default(&flow_params);
        variational(wx, wy, seq[0], seq[1], &flow_params);
        return (int)r.x + (occlusions != 0);
    }
    sf_mt_params_t m;
    sf_mt_params_from_list(thread_params, &m);
    return (m.S == 3 && m.robust_grad == -1 && m.niter_alter == 1) ? 0 : 1;
}
