/* Prints sizeof / offsetof of every field of the C ABI's POD structs (include/melspec_b200.h), one "struct.field offset size"
 * line each.  tests/test_layout.py compares this with the ctypes mirror (mel-spec_b200/_lib.py) and with the `// offset N`
 * comments of the Rust `#[repr(C)]` structs (rust/src/ffi.rs), so that a reordered or retyped field fails the test-suite. */
#include <stddef.h>
#include <stdio.h>

#include "melspec_b200.h"

#define F(S, f) printf(#S "." #f " %zu %zu\n", offsetof(S, f), sizeof(((S*)0)->f))

int main(void) {
    printf("melspec_config.sizeof %zu 0\n", sizeof(melspec_config));
    F(melspec_config, frontend); F(melspec_config, fft_size); F(melspec_config, hop_size); F(melspec_config, n_mels);
    F(melspec_config, sampling_rate); F(melspec_config, frame_length); F(melspec_config, apply_cmn);
    F(melspec_config, use_log_fbank); F(melspec_config, use_power); F(melspec_config, preemphasis); F(melspec_config, low_freq);
    F(melspec_config, high_freq); F(melspec_config, energy_floor); F(melspec_config, win_length); F(melspec_config, center);
    F(melspec_config, pad_to); F(melspec_config, normalize_per_feature); F(melspec_config, htk); F(melspec_config, slaney_norm);
    F(melspec_config, log_zero_guard); F(melspec_config, f_min); F(melspec_config, f_max);
    printf("melspec_vad_settings.sizeof %zu 0\n", sizeof(melspec_vad_settings));
    F(melspec_vad_settings, min_energy); F(melspec_vad_settings, min_y); F(melspec_vad_settings, min_x); F(melspec_vad_settings, min_mel);
    printf("abi %d 0\n", MELSPEC_B200_ABI_VERSION);
    return 0;
}
