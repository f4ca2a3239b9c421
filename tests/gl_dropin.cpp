// Compiles the PBF_WITH_GL branch of the shims (include/pbf/SPH.h) the way the reference's tree would -- "common.h" is
// the reference's src/common.h there, tests/gl_stub/common.h here -- and replays every call the reference's Simulation
// makes on its SPH member (file:line of /root/reference/src/Simulation.cpp at each site).  Without a GL context the
// registration inside SPH::SPH fails with a std::runtime_error that names the cause; the test expects exactly that
// (a display-less box can go no further), which proves the branch compiles, links and fails cleanly.
#include "common.h"          // src/Simulation.h:25
#include "pbf/SPH.h"         // replaces src/SPH.h (src/Simulation.h:32)

#include <vector>

struct PointSpriteStub {     // src/PointSprite.cpp:64-84: binds the buffer names as vertex attributes
    GLuint pos = 0, hl = 0;
    void SetPositionBuffer(GLuint buffer, GLsizeiptr stride, GLintptr offset) { pos = buffer; (void)stride; (void)offset; }
    void SetHighlightBuffer(GLuint buffer, GLsizeiptr stride, GLintptr offset) { hl = buffer; (void)stride; (void)offset; }
};

class SimulationCallSites {
public:
    SimulationCallSites() : sph(GetNumberOfParticles()) {                       // src/Simulation.cpp:25-28
        ResetParticleBuffer();                                                   // :75
        pointsprite.SetPositionBuffer(sph.GetPositionBuffer(), 4 * sizeof(float), 0);   // :78
        pointsprite.SetHighlightBuffer(sph.GetHighlightBuffer(), sizeof(GLuint), 0);    // :79
    }
    unsigned int GetNumberOfParticles() const { return 32 * 32 * 32 * 2; }       // :200-204
    void ResetParticleBuffer() {                                                 // :249-272
        std::vector<float> positions(4 * (size_t)GetNumberOfParticles(), 1.0f), velocities(positions.size(), 0.0f);
        GLuint tmpbuffer;
        glGenBuffers(1, &tmpbuffer);
        glBindBuffer(GL_COPY_READ_BUFFER, tmpbuffer);
        glBufferData(GL_COPY_READ_BUFFER, sizeof(float) * positions.size(), &positions[0], GL_STREAM_COPY);
        glBindBuffer(GL_COPY_WRITE_BUFFER, sph.GetPositionBuffer());
        glCopyBufferSubData(GL_COPY_READ_BUFFER, GL_COPY_WRITE_BUFFER, 0, 0, sizeof(float) * positions.size());
        glBufferData(GL_COPY_READ_BUFFER, sizeof(float) * velocities.size(), &velocities[0], GL_STREAM_COPY);
        glBindBuffer(GL_COPY_WRITE_BUFFER, sph.GetVelocityBuffer());
        glCopyBufferSubData(GL_COPY_READ_BUFFER, GL_COPY_WRITE_BUFFER, 0, 0, sizeof(float) * velocities.size());
        glDeleteBuffers(1, &tmpbuffer);
        glBindBuffer(GL_SHADER_STORAGE_BUFFER, sph.GetHighlightBuffer());
        glClearBufferData(GL_SHADER_STORAGE_BUFFER, GL_R8UI, GL_RED_INTEGER, GL_UNSIGNED_INT, NULL);
    }
    void ToggleHighlight(GLint id) {                                             // :160-195
        GLuint tmpbuffer;
        glGenBuffers(1, &tmpbuffer);
        glBindBuffer(GL_COPY_WRITE_BUFFER, tmpbuffer);
        glBufferData(GL_COPY_WRITE_BUFFER, sizeof(GLuint), NULL, GL_DYNAMIC_READ);
        glBindBuffer(GL_COPY_READ_BUFFER, sph.GetHighlightBuffer());
        glCopyBufferSubData(GL_COPY_READ_BUFFER, GL_COPY_WRITE_BUFFER, id * sizeof(GLuint), 0, sizeof(GLuint));
        GLuint *info = reinterpret_cast<GLuint *>(glMapBuffer(GL_COPY_WRITE_BUFFER, GL_READ_WRITE));
        if (info == NULL) throw std::runtime_error("A GPU buffer could not be mapped to CPU address space.");
        *info = *info > 0 ? 0 : 1;
        glUnmapBuffer(GL_COPY_WRITE_BUFFER);
        glCopyBufferSubData(GL_COPY_WRITE_BUFFER, GL_COPY_READ_BUFFER, 0, id * sizeof(GLuint), sizeof(GLuint));
        glDeleteBuffers(1, &tmpbuffer);
    }
    void Frame(bool running) {                                                   // :464-479
        if (running) sph.Run();
        pointsprite.SetPositionBuffer(sph.GetPositionBuffer(), 4 * sizeof(float), 0);
    }
    void Keys() {                                                                // :280-282, :299-318, :348-350, :389-420
        sph.SetExternalForce(true);
        sph.SetExternalForce(false);
        sph.SetVorticityConfinementEnabled(!sph.IsVorticityConfinementEnabled());
        sph.OutputTiming();
        sph.Run();
        sph.SetRestDensity(sph.GetRestDensity() + 0.01f);
        sph.SetCFMEpsilon(sph.GetCFMEpsilon() + 1.0f);
        sph.SetGravity(sph.GetGravity());
        sph.SetTimestep(sph.GetTimestep() + 0.001f);
        sph.SetNumSolverIterations(sph.GetNumSolverIterations() + 1);
        sph.SetTensileInstabilityK(sph.GetTensileInstabilityK() + 0.1f);
        sph.SetTensileInstabilityScale(1.0f / SPH::Wpoly6(0.2f, 2.0f));
        sph.SetXSPHViscosity(sph.GetXSPHViscosity() + 0.01f);
        sph.SetVorticityEpsilon(sph.GetVorticityEpsilon() + 0.1f);
    }

private:
    SPH sph;
    PointSpriteStub pointsprite;
};

int main() {
    try {
        SimulationCallSites sim;
        sim.ToggleHighlight(5);
        sim.Frame(true);
        sim.Keys();
        printf("GLDROPIN ran with a GL context\n");
    } catch (const std::runtime_error &e) {
        // no device, or a device but no GL context: both are clean failures of the constructor
        printf("GLDROPIN expected error: %s (gl stub calls: %d)\n", e.what(), gl_stub::state().calls);
    }
    return 0;
}
