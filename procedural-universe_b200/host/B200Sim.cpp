#include "B200Sim.hpp"

#include "Core/Event.hpp"
#include "Services/Log.hpp"
#include "Sim/Octree.hpp"

#include <string>

static_assert(sizeof(Particle) == NB_PARTICLE_STRIDE, "Particle layout differs from the one the engine was built for");
static_assert(offsetof(Particle, Velocity) == NB_OFF_VELOCITY && offsetof(Particle, Forces) == NB_OFF_FORCES &&
              offsetof(Particle, Mass) == NB_OFF_MASS, "Particle layout differs from the one the engine was built for");

B200Sim::B200Sim(ID3D11DeviceContext* context, EMode mode, int device) : Mode(mode)
{
    LOGM(mode == EMode::AllPairs ? "Brute Force B200" : "Barnes-Hut B200")

    nb_config cfg;
    nb_default_config(&cfg);
    cfg.device = device;
    cfg.mode = (mode == EMode::AllPairs) ? NB_MODE_ALLPAIRS : NB_MODE_BARNESHUT;
    cfg.theta = static_cast<float>(Octree::Theta);          // process-global in the reference (Octree.cpp:5)
    if (nb_create(&cfg, &Handle) != NB_OK)
    {
        LOGE(std::string("B200 engine unavailable: ") + nb_last_error())
        Handle = nullptr;                                    // Update() becomes a logged no-op, like
        return;                                              // BruteForceGPU without its shader (:25-33)
    }

    if (context && mode == EMode::BarnesHut)
        DebugCube = std::make_unique<Cube>(context);         // BarnesHut.cpp:23-27

    if (mode == EMode::BarnesHut)
    {
        // BarnesHut.cpp:29-31: the handler writes the process-global Octree::Theta (so sims created
        // later start from it) and this sim follows.
        EventStream::Register(EEvent::BHThetaChanged, [this](const EventData& data) {
            Octree::Theta = EventValue<FloatEventData>(data);
            SetTheta(static_cast<float>(Octree::Theta));
        });
        ThetaRegistered = true;
    }
}

B200Sim::~B200Sim()
{
    Shutdown();
}

void B200Sim::Shutdown()
{
    // BarnesHut::~BarnesHut (BarnesHut.cpp:34-37): drop the BHThetaChanged handler, which captures `this`
    if (ThetaRegistered)
    {
        EventStream::UnregisterAll(EEvent::BHThetaChanged);
        ThetaRegistered = false;
    }
    Unpin();
    if (Handle)
    {
        nb_destroy(Handle);
        Handle = nullptr;
    }
}

void B200Sim::SetTheta(float theta)
{
    if (Handle && nb_set_theta(Handle, theta) != NB_OK)
        LOGE(std::string("B200Sim::SetTheta: ") + nb_last_error())
}

void B200Sim::Pin()
{
    Unpin();
    if (!Particles || Particles->empty()) return;
    // Page-lock the caller's array so the per-Update upload / write-back run at full PCIe rate.
    if (nb_host_register(Particles->data(), Particles->size() * sizeof(Particle)) == NB_OK)
    {
        Pinned = Particles->data();
        PinnedBytes = Particles->size() * sizeof(Particle);
    }
}

void B200Sim::Unpin()
{
    if (Pinned) nb_host_unregister(Pinned);
    Pinned = nullptr;
    PinnedBytes = 0;
}

void B200Sim::Init(std::vector<Particle>& particles)
{
    Particles = &particles;
    if (!Handle || particles.empty()) return;
    Pin();
    if (nb_init_aos(Handle, particles.data(), particles.size(), sizeof(Particle)) != NB_OK)
        LOGE(std::string("B200Sim::Init: ") + nb_last_error())
}

void B200Sim::Update(float dt)
{
    if (!Handle || !Particles || Particles->empty()) return;
    // the vector may have been reallocated by the caller since Init (resize + re-Init is the
    // reference's protocol, SimulationState.cpp:218-227, but stay safe if it was not followed)
    if (Pinned != Particles->data() || PinnedBytes != Particles->size() * sizeof(Particle)) Pin();
    if (nb_update_aos(Handle, Particles->data(), Particles->size(), sizeof(Particle), dt) != NB_OK)
        LOGE(std::string("B200Sim::Update: ") + nb_last_error())
}

void B200Sim::RenderDebug(DirectX::SimpleMath::Matrix view, DirectX::SimpleMath::Matrix proj)
{
    if (!Handle || !DebugCube || Mode != EMode::BarnesHut) return;
    size_t n = 0;
    if (nb_get_leaf_cells(Handle, nullptr, nullptr, &n) != NB_OK || n == 0) return;
    DebugCells.resize(4 * n);
    if (nb_get_leaf_cells(Handle, DebugCells.data(), nullptr, &n) != NB_OK)
    {
        LOGE(std::string("B200Sim::RenderDebug: ") + nb_last_error())
        return;
    }
    for (size_t k = 0; k < n; ++k)      // Octree::RenderDebug: cube->Render(pos, size, view * proj), Octree.cpp:166
        DebugCube->Render(DirectX::SimpleMath::Vector3(DebugCells[4 * k], DebugCells[4 * k + 1], DebugCells[4 * k + 2]),
                          DebugCells[4 * k + 3], view * proj);
}

std::unique_ptr<INBodySim> CreateB200NBodySim(ID3D11DeviceContext* context, ENBodySim type)
{
    switch (type)
    {
        case ENBodySim::BruteForceGPU:
            return std::make_unique<B200Sim>(context, B200Sim::EMode::AllPairs);
        case ENBodySim::BarnesHut:
            return std::make_unique<B200Sim>(context, B200Sim::EMode::BarnesHut);
        default:
            return CreateNBodySim(context, type);
    }
}
