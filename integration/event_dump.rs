// event_dump.rs -- instrumentation module for the Rust reference (bacpop/Pansim).
// SOURCE ONLY: no Rust toolchain exists in the image this repository is built in, so this
// file has not been compiled. It records, per generation, everything the apply step of
// pansim/src/main.rs:445-464 consumes and writes it in the "PSEV" format that
// pansim_b200/event_dump.py reads and `pansim_step_replay` (include/pansim_b200.h) applies.
//
// How to wire it in (pansim/src/):
//   lib.rs            add `pub mod event_dump;`
//   population.rs     `use crate::event_dump::EventDump;` and give `mutate_alleles` and
//                     `recombine` an extra `dump: Option<&std::sync::Mutex<EventDump>>` argument
//     :501-509 / :525-538  inside the per-row closure collect `(mutant_site, new_allele)` in a
//                     local Vec and, after the loop, `dump.lock().push_row_mutations(core, row_idx, &local)`
//                     (use `.enumerate()` on the `axis_iter_mut` to get `row_idx`); rows are
//                     disjoint, so the order in which rows are pushed does not matter --
//                     `finish_generation` sorts the per-row blocks by row.
//     :741-746        inside the serial apply loop: `dump.push_transfer(core, row_idx, col_idx, value)`
//   main.rs :443      after `sample_indices`: `dump.set_parents(&sampled_individuals)`
//   main.rs :464      after the two `recombine` calls: `dump.finish_generation(j as u32, &mut file)?`
use std::fs::File;
use std::io::{self, Write};

#[derive(Default)]
pub struct EventDump {
    parents: Vec<u32>,
    core_rows: Vec<(u32, Vec<(u32, u8)>)>, // (row, [(site, one-hot allele)]) in draw order
    acc_rows: Vec<(u32, Vec<u32>)>,        // (row, [gene]) flips
    hr: Vec<(u32, u32, u8)>,               // (recipient, locus, value) in apply order
    hgt: Vec<(u32, u32)>,                  // (recipient, gene) in apply order
}

impl EventDump {
    pub fn set_parents(&mut self, parents: &[usize]) {
        self.parents = parents.iter().map(|&p| p as u32).collect();
    }

    /// population.rs:501-509 (accessory, alleles ignored) and :525-538 (core)
    pub fn push_row_mutations(&mut self, core: bool, row: usize, events: &[(usize, u8)]) {
        if core {
            self.core_rows.push((row as u32, events.iter().map(|&(s, a)| (s as u32, a)).collect()));
        } else {
            self.acc_rows.push((row as u32, events.iter().map(|&(s, _)| s as u32).collect()));
        }
    }

    /// population.rs:741-746, called in the order the reference performs the stores
    pub fn push_transfer(&mut self, core: bool, recipient: usize, locus: usize, value: u8) {
        if core {
            self.hr.push((recipient as u32, locus as u32, value));
        } else {
            self.hgt.push((recipient as u32, locus as u32));
        }
    }

    pub fn finish_generation(&mut self, gen: u32, out: &mut File) -> io::Result<()> {
        self.core_rows.sort_by_key(|r| r.0);
        // accessory compartments are mutated one after the other (population.rs:476); a stable
        // sort keeps compartment order inside a row, and flips commute anyway
        self.acc_rows.sort_by_key(|r| r.0);
        let n_core: u64 = self.core_rows.iter().map(|r| r.1.len() as u64).sum();
        let n_acc: u64 = self.acc_rows.iter().map(|r| r.1.len() as u64).sum();
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
        for (row, ev) in &self.core_rows { for _ in ev { out.write_all(&row.to_le_bytes())?; } }
        for (_, ev) in &self.core_rows { for (s, _) in ev { out.write_all(&s.to_le_bytes())?; } }
        for (_, ev) in &self.core_rows { for (_, a) in ev { out.write_all(&[*a])?; } }
        for (row, ev) in &self.acc_rows { for _ in ev { out.write_all(&row.to_le_bytes())?; } }
        for (_, ev) in &self.acc_rows { for g in ev { out.write_all(&g.to_le_bytes())?; } }
        for (r, _, _) in &self.hr { out.write_all(&r.to_le_bytes())?; }
        for (_, l, _) in &self.hr { out.write_all(&l.to_le_bytes())?; }
        for (_, _, v) in &self.hr { out.write_all(&[*v])?; }
        for (r, _) in &self.hgt { out.write_all(&r.to_le_bytes())?; }
        for (_, g) in &self.hgt { out.write_all(&g.to_le_bytes())?; }
        *self = EventDump::default();
        Ok(())
    }
}
