// event_dump.rs -- instrumentation module for the Rust reference (bacpop/Pansim).
//
// SOURCE ONLY: no Rust toolchain exists in the image this repository is built in, so this
// file has not been compiled here. `integration/instrument_population.patch` adds it to the
// reference as `pansim/src/event_dump.rs` together with the hooks listed below; a CPU test
// (`tests/test_instrumentation_patch.py`) checks that the patch applies to the reference tree.
//
// With `PANSIM_EVENT_DUMP=<file>` in the environment the patched `pansim` binary records, per
// generation, everything the apply step of pansim/src/main.rs:445-464 consumes and appends it to
// <file> in the "PSEV" format that `pansim_b200/event_dump.py` reads and `pansim_step_replay`
// (include/pansim_b200.h) applies bit for bit. Without the variable every hook is one relaxed
// atomic load.
//
// Hooks (line numbers of the unpatched pansim/src/population.rs):
//   :443       `set_parents(&sampled_indices)`            the N parent draws of sample_indices
//   :501-509   `push_row_mutations(false, row, &events)`  accessory flips of one row, draw order
//   :525-538   `push_row_mutations(true, row, &events)`   core SNPs of one row, draw order
//   :741-746   `push_transfer(core, donor, recipient, locus, value)`   inside the serial apply
//              loop, i.e. in the order the reference performs `self.pop[[row, col]] = value`
//   main.rs:464 `finish_generation(j)`                     after the two recombine calls
//
// The mutation hooks run under rayon: rows are disjoint, so the order in which rows arrive is
// irrelevant; `finish_generation` sorts the per-row blocks by row (stable, so the compartment
// order of population.rs:476 is kept inside a row).
use std::fs::File;
use std::io::{self, BufWriter, Write};
use std::sync::atomic::{AtomicBool, Ordering};
use std::sync::Mutex;

static ENABLED: AtomicBool = AtomicBool::new(false);
static STATE: Mutex<Option<EventDump>> = Mutex::new(None); // const Mutex::new needs Rust >= 1.63

struct EventDump {
    out: BufWriter<File>,
    parents: Vec<u32>,
    core_rows: Vec<(u32, Vec<(u32, u8)>)>, // (row, [(site, one-hot allele)]) in draw order
    acc_rows: Vec<(u32, Vec<u32>)>,        // (row, [gene]) flips
    hr: Vec<(u32, u32, u8)>,               // (recipient, locus, value) in apply order
    hgt: Vec<(u32, u32)>,                  // (recipient, gene) in apply order
}

/// `PANSIM_EVENT_DUMP=<file>` switches the recorder on (called once from main.rs).
pub fn init_from_env() {
    if let Ok(path) = std::env::var("PANSIM_EVENT_DUMP") {
        let file = File::create(&path).expect("PANSIM_EVENT_DUMP: cannot create the dump file");
        *STATE.lock().unwrap() = Some(EventDump {
            out: BufWriter::new(file),
            parents: Vec::new(),
            core_rows: Vec::new(),
            acc_rows: Vec::new(),
            hr: Vec::new(),
            hgt: Vec::new(),
        });
        ENABLED.store(true, Ordering::Relaxed);
    }
}

#[inline]
pub fn enabled() -> bool {
    ENABLED.load(Ordering::Relaxed)
}

/// population.rs:443
pub fn set_parents(parents: &[usize]) {
    if !enabled() {
        return;
    }
    if let Some(d) = STATE.lock().unwrap().as_mut() {
        d.parents = parents.iter().map(|&p| p as u32).collect();
    }
}

/// population.rs:501-509 (accessory; the allele is implied by the flip) and :525-538 (core)
pub fn push_row_mutations(core: bool, row: usize, events: &[(usize, u8)]) {
    if !enabled() || events.is_empty() {
        return;
    }
    if let Some(d) = STATE.lock().unwrap().as_mut() {
        if core {
            d.core_rows
                .push((row as u32, events.iter().map(|&(s, a)| (s as u32, a)).collect()));
        } else {
            d.acc_rows
                .push((row as u32, events.iter().map(|&(s, _)| s as u32).collect()));
        }
    }
}

/// population.rs:741-746, called in the order the reference performs the stores
pub fn push_transfer(core: bool, _donor: usize, recipient: usize, locus: usize, value: u8) {
    if !enabled() {
        return;
    }
    if let Some(d) = STATE.lock().unwrap().as_mut() {
        if core {
            d.hr.push((recipient as u32, locus as u32, value));
        } else {
            d.hgt.push((recipient as u32, locus as u32));
        }
    }
}

/// main.rs:464: one PSEV record per generation
pub fn finish_generation(gen: u32) {
    if !enabled() {
        return;
    }
    if let Some(d) = STATE.lock().unwrap().as_mut() {
        d.write_record(gen).expect("PANSIM_EVENT_DUMP: write failed");
    }
}

impl EventDump {
    fn write_record(&mut self, gen: u32) -> io::Result<()> {
        self.core_rows.sort_by_key(|r| r.0);
        self.acc_rows.sort_by_key(|r| r.0);
        let n_core: u64 = self.core_rows.iter().map(|r| r.1.len() as u64).sum();
        let n_acc: u64 = self.acc_rows.iter().map(|r| r.1.len() as u64).sum();
        let out = &mut self.out;
        out.write_all(b"PSEV")?;
        for v in [1u32, gen, self.parents.len() as u32] {
            out.write_all(&v.to_le_bytes())?;
        }
        for v in [n_core, n_acc, self.hr.len() as u64, self.hgt.len() as u64] {
            out.write_all(&v.to_le_bytes())?;
        }
        for p in &self.parents {
            out.write_all(&p.to_le_bytes())?;
        }
        for (row, ev) in &self.core_rows {
            for _ in ev {
                out.write_all(&row.to_le_bytes())?;
            }
        }
        for (_, ev) in &self.core_rows {
            for (s, _) in ev {
                out.write_all(&s.to_le_bytes())?;
            }
        }
        for (_, ev) in &self.core_rows {
            for (_, a) in ev {
                out.write_all(&[*a])?;
            }
        }
        for (row, ev) in &self.acc_rows {
            for _ in ev {
                out.write_all(&row.to_le_bytes())?;
            }
        }
        for (_, ev) in &self.acc_rows {
            for g in ev {
                out.write_all(&g.to_le_bytes())?;
            }
        }
        for (r, _, _) in &self.hr {
            out.write_all(&r.to_le_bytes())?;
        }
        for (_, l, _) in &self.hr {
            out.write_all(&l.to_le_bytes())?;
        }
        for (_, _, v) in &self.hr {
            out.write_all(&[*v])?;
        }
        for (r, _) in &self.hgt {
            out.write_all(&r.to_le_bytes())?;
        }
        for (_, g) in &self.hgt {
            out.write_all(&g.to_le_bytes())?;
        }
        out.flush()?;
        self.parents.clear();
        self.core_rows.clear();
        self.acc_rows.clear();
        self.hr.clear();
        self.hgt.clear();
        Ok(())
    }
}
