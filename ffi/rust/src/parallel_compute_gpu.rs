// Replacement for the body of calculate_eigenvalues_parallel (src/data_storage/parallel_compute.rs:14-41):
// one batched GPU call per chunk instead of one rayon task per seed, the next chunk in flight while this
// chunk's rows go to the writer thread.  Signature and channel protocol are the reference's.
use std::sync::mpsc;

use crate::gpu_ffi::Gpu;
use crate::johansen_models::JohansenModel;

const GPU_CHUNK: usize = 1 << 20; // replaces BATCH_SIZE = 10_000 (src/data_storage/config.rs)

pub(crate) fn calculate_eigenvalues_parallel(
    dim: usize,
    steps: usize,
    seeds: &[u32],
    model: JohansenModel,
    sender: mpsc::Sender<(u32, Vec<f64>)>,
    quiet: bool,
) {
    // the reference panics inside the hot path (johansen_statistics.rs:45, :192); keep that policy
    let gpu = Gpu::new().unwrap_or_else(|e| panic!("GPU init failed: {e}"));
    let m = model.to_number();
    let p = Gpu::num_eigs(m, dim);
    let mut chunks = seeds.chunks(GPU_CHUNK);
    let mut current = chunks.next();
    let mut pending = current.map(|c| gpu.submit(m, dim, steps, c).unwrap_or_else(|e| panic!("GPU submit failed: {e}")));
    while let (Some(chunk), Some(batch)) = (current, pending.take()) {
        let rows = gpu.wait(batch).unwrap_or_else(|e| panic!("GPU batch failed: {e}"));
        current = chunks.next();
        pending = current.map(|c| gpu.submit(m, dim, steps, c).unwrap_or_else(|e| panic!("GPU submit failed: {e}")));
        for (i, &seed) in chunk.iter().enumerate() {
            if sender.send((seed, rows[i * p..(i + 1) * p].to_vec())).is_err() && !quiet {
                eprintln!("Failed to send results to writer thread");
            }
        }
    }
}

// The same through the streaming entry point: no chunk loop and no intermediate Vec of rows on this side -- the
// library hands every finished range to the closure from its own host threads, and the closure forwards the records
// to the writer thread exactly as the rayon tasks of the reference do (`mpsc::Sender` is `Send`; one clone per call
// keeps the closure `Sync`).
pub(crate) fn calculate_eigenvalues_parallel_streaming(
    dim: usize,
    steps: usize,
    seeds: &[u32],
    model: JohansenModel,
    sender: mpsc::Sender<(u32, Vec<f64>)>,
    quiet: bool,
) {
    let gpu = Gpu::new().unwrap_or_else(|e| panic!("GPU init failed: {e}"));
    let m = model.to_number();
    let p = Gpu::num_eigs(m, dim);
    let sender = std::sync::Mutex::new(sender);
    gpu.eigs_batch_multi_stream(1u32 << m, dim, steps, seeds, |first, rows| {
        let tx = sender.lock().unwrap().clone();
        for (i, row) in rows.chunks_exact(p).enumerate() {
            if tx.send((seeds[first + i], row.to_vec())).is_err() && !quiet {
                eprintln!("Failed to send results to writer thread");
            }
        }
        true
    })
    .unwrap_or_else(|e| panic!("GPU batch failed: {e}"));
}
