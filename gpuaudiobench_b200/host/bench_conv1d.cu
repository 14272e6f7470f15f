#include "bench_conv1d.cuh"

#include <cstdio>
#include <stdexcept>
#include <vector>

#include "benchmark_constants.cuh"

Conv1DBenchmark::Conv1DBenchmark(int ir_length, size_t buffer_size, size_t track_count)
    : GPUABenchmark("Conv1D", buffer_size, track_count), ir_length_(ir_length) {}

Conv1DBenchmark::~Conv1DBenchmark() {
    BenchmarkUtils::freeHostBuffers({h_ir_buf, cpu_reference});
    h_ir_buf = cpu_reference = nullptr;
}

void Conv1DBenchmark::allocateConvBuffers() {
    h_ir_buf = BenchmarkUtils::allocateHostBuffer<float>(getTrackCount() * ir_length_, benchmark_name_ + " host IR buffer");
    cpu_reference = BenchmarkUtils::allocateHostBuffer<float>(getTotalElements(), "conv1d cpu reference");
}

void Conv1DBenchmark::generateImpulseResponses() {
    ConvCommon::generateImpulseResponses(h_ir_buf, getTrackCount(), ir_length_, ConvCommon::IRVariant::DIRECT_FLOAT_PI);
}

// The reference's kernel and CPU loop bound the FLAT input index (bench_conv1d.cu:17-24,197-199),
// so track t "remembers" the tail of the tracks before it.  A streaming engine gets the same
// numbers when track t's history is primed with x_flat[t*B-(L-1) .. t*B-1] (zero before index 0).
void Conv1DBenchmark::primeReferenceHistory() {
    const size_t T = getTrackCount(), B = getBufferSize();
    const size_t H = static_cast<size_t>(ir_length_) - 1;
    auto prime = [&](const float* h) {
        if (group_.valid()) group_.primeHistory(h); else engine_.primeHistory(h);
    };
    if (H == 0) {
        prime(nullptr);
        return;
    }
    std::vector<float>& hist = history_;
    hist.assign(T * H, 0.0f);
    const float* x = getHostInput();
    for (size_t t = 0; t < T; ++t) {
        const long long first = static_cast<long long>(t * B) - static_cast<long long>(H);  // flat index of hist[t][0]
        for (size_t i = 0; i < H; ++i) {
            const long long idx = first + static_cast<long long>(i);
            if (idx >= 0) hist[t * H + i] = x[idx];
        }
    }
    prime(hist.data());
}

void Conv1DBenchmark::calculateCPUReference() {
    ConvCommon::cpuConvFlatHistory(getHostInput(), h_ir_buf, cpu_reference, ir_length_, static_cast<int>(getBufferSize()),
                                   static_cast<int>(getTrackCount()));
}

void Conv1DBenchmark::setupBenchmark() {
    allocateBuffers(getTotalElements());
    allocateConvBuffers();
    generateImpulseResponses();
    if (NGPUS > 1) {
        group_.create(B200CONV_ALGO_DIRECT, B200CONV_OUT_TRACK_MAJOR, getTrackCount(), getBufferSize(), ir_length_, NGPUS);
        group_.loadIR(h_ir_buf);
    } else {
        engine_.create(B200CONV_ALGO_DIRECT, B200CONV_OUT_TRACK_MAJOR, getTrackCount(), getBufferSize(), ir_length_);
        engine_.loadIR(h_ir_buf);
    }
    generateTestData(42);
    primeReferenceHistory();
    calculateCPUReference();
    ready_ = true;
    std::printf("Conv1D benchmark setup complete (IR length = %d, B200 direct-form engine, %d GPU%s, %s mode)\n", ir_length_,
                NGPUS, NGPUS > 1 ? "s" : "", STREAM_MODE ? "streaming" : "stateless");
}

void Conv1DBenchmark::oneIteration(const char* caller) {
    if (!ready_) throw std::runtime_error(std::string("Conv1DBenchmark::") + caller + " called before setupBenchmark");
    if (group_.valid()) {  // tracks sharded over NGPUS devices: host buffers in, host buffers out
        group_.processHost(getHostInput(), getHostOutput(), nullptr, /*advance_state=*/STREAM_MODE);
        return;
    }
    transferToDevice();
    BenchmarkUtils::CudaEventTimer gpu;
    gpu.start();
    // stateless by default: the reference re-submits the same buffer every iteration (bench_base.cu:89-94)
    engine_.process(getDeviceInput(), getDeviceOutput(), nullptr, /*advance_state=*/STREAM_MODE, nullptr);
    recordGpuDuration(gpu.stop());
    transferToHost();
}

void Conv1DBenchmark::runKernel() { oneIteration("runKernel"); }
void Conv1DBenchmark::performBenchmarkIteration() { oneIteration("performBenchmarkIteration"); }

void Conv1DBenchmark::validate(ValidationData& validation_data) {
    using namespace BenchmarkConstants;
    if (STREAM_MODE) {  // the timed loop advanced the stream: recreate the state the CPU reference describes
        primeReferenceHistory();
        if (group_.valid()) {
            group_.processHost(getHostInput(), getHostOutput(), nullptr, false);
        } else {
            transferToDevice();
            engine_.process(getDeviceInput(), getDeviceOutput(), nullptr, false, nullptr);
            synchronizeAndCheck();
            transferToHost();
        }
    }
    validation_data = compareWithReference(cpu_reference, CONV1D_REFERENCE_ABS_TOL);  // the reference's check
    const ConvCommon::Accuracy acc = ConvCommon::measureAccuracy(getHostOutput(), cpu_reference, getTotalElements());
    const bool stated_ok = acc.snr_db >= CONV1D_MIN_SNR_DB && acc.max_abs_err <= CONV1D_MAX_ABS_REL_TO_PEAK * acc.ref_peak;
    validation_data.messages.push_back(ConvCommon::describeAccuracy(acc, CONV1D_MIN_SNR_DB, CONV1D_MAX_ABS_REL_TO_PEAK));
    if (!stated_ok && validation_data.status == ValidationStatus::SUCCESS) {
        validation_data.status = ValidationStatus::FAILURE;
        validation_data.messages.insert(validation_data.messages.begin(), "Validation failed: stated fp32 tolerance exceeded");
    }
    if (validation_data.status == ValidationStatus::SUCCESS) validation_data.messages.push_back("Conv1D validation passed");
}
