! iso_c_binding shim: multicharge's model API on top of libmchrg_b200.so (include/mchrg_b200.h).
!
! SOURCE ONLY -- this image has no Fortran compiler and mctc-lib v0.5.2 is not vendored, so this
! file is neither compiled nor tested here (see INTEGRATION.md).  It is the binding a maintainer
! adds to the reference tree: `b200_model_type` extends `mchrg_model_type`, overrides `solve`
! (src/multicharge/model/type.F90:151-292) and implements the deferred procedures
! (type.F90:78-125) by forwarding the caller's arrays -- unchanged, column-major -- to the C ABI.
module mchrg_b200_shim
   use, intrinsic :: iso_c_binding
   use mctc_env, only : wp, error_type, fatal_error
   use mctc_io, only : structure_type
   use multicharge_model_type, only : mchrg_model_type
   use multicharge_model_cache, only : cache_container
   implicit none
   private

   public :: b200_model_type, new_b200_eeq2019_model, new_b200_eeqbc2025_model, b200_get_charges
   public :: b200_get_charges_batch, b200_comm_init_nccl, b200_comm_unique_id

   type, bind(C) :: c_structure
      integer(c_int32_t) :: nat
      integer(c_int32_t) :: nid
      type(c_ptr) :: id
      type(c_ptr) :: xyz
      real(c_double) :: charge
      integer(c_int32_t) :: periodic(3)
      real(c_double) :: lattice(9)
   end type c_structure

   !> multicharge model whose hot path runs on a B200 (drop-in for eeq_model / eeqbc_model)
   type, extends(mchrg_model_type) :: b200_model_type
      type(c_ptr) :: ctx = c_null_ptr
      type(c_ptr) :: handle = c_null_ptr
   contains
      procedure :: solve => b200_solve
      procedure :: update => b200_update
      procedure :: get_xvec => b200_get_xvec
      procedure :: get_xvec_derivs => b200_get_xvec_derivs
      procedure :: get_coulomb_matrix => b200_get_coulomb_matrix
      procedure :: get_coulomb_derivs => b200_get_coulomb_derivs
      final :: b200_free
   end type b200_model_type

   interface
      function c_context_create(device, ctx) result(stat) bind(C, name="mchrg_b200_context_create")
         import :: c_int, c_ptr
         integer(c_int), value :: device
         type(c_ptr), intent(out) :: ctx
         integer(c_int) :: stat
      end function
      function c_new_eeq2019_model(ctx, nid, num, model) result(stat) bind(C, name="mchrg_b200_new_eeq2019_model")
         import :: c_int, c_int32_t, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int32_t), value :: nid
         integer(c_int32_t), intent(in) :: num(*)
         type(c_ptr), intent(out) :: model
         integer(c_int) :: stat
      end function
      function c_new_eeqbc2025_model(ctx, nid, num, rvdw, en, model) result(stat) &
            & bind(C, name="mchrg_b200_new_eeqbc2025_model")
         import :: c_int, c_int32_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int32_t), value :: nid
         integer(c_int32_t), intent(in) :: num(*)
         real(c_double), intent(in) :: rvdw(*), en(*)   ! get_vdw_rad(pair)*autoaa per species pair, get_pauling_en
         type(c_ptr), intent(out) :: model
         integer(c_int) :: stat
      end function
      function c_local_charge(model, mol, qloc, dqlocdr, dqlocdL) result(stat) bind(C, name="mchrg_b200_local_charge")
         import :: c_int, c_ptr, c_structure
         type(c_ptr), value :: model
         type(c_structure), intent(in) :: mol
         type(c_ptr), value :: qloc, dqlocdr, dqlocdL
         integer(c_int) :: stat
      end function
      function c_singlepoint_batch(model, nmol, offsets, num, xyz, charge, qvec, energy, gradient, status) result(stat) &
            & bind(C, name="mchrg_b200_singlepoint_batch")
         import :: c_int, c_int32_t, c_ptr
         type(c_ptr), value :: model
         integer(c_int32_t), value :: nmol
         type(c_ptr), value :: offsets, num, xyz, charge, qvec, energy, gradient, status
         integer(c_int) :: stat
      end function
      subroutine c_model_free(model) bind(C, name="mchrg_b200_model_free")
         import :: c_ptr
         type(c_ptr), value :: model
      end subroutine
      function c_last_error() result(msg) bind(C, name="mchrg_b200_last_error")
         import :: c_ptr
         type(c_ptr) :: msg
      end function
      function c_solve(model, mol, cn, qloc, dcndr, dcndL, dqlocdr, dqlocdL, energy, gradient, sigma, &
            & qvec, dqdr, dqdL) result(stat) bind(C, name="mchrg_b200_solve")
         import :: c_int, c_ptr, c_structure
         type(c_ptr), value :: model
         type(c_structure), intent(in) :: mol
         type(c_ptr), value :: cn, qloc, dcndr, dcndL, dqlocdr, dqlocdL
         type(c_ptr), value :: energy, gradient, sigma, qvec, dqdr, dqdL
         integer(c_int) :: stat
      end function
      function c_get_charges(model, mol, qvec, dqdr, dqdL) result(stat) bind(C, name="mchrg_b200_get_charges")
         import :: c_int, c_ptr, c_structure
         type(c_ptr), value :: model
         type(c_structure), intent(in) :: mol
         type(c_ptr), value :: qvec, dqdr, dqdL
         integer(c_int) :: stat
      end function
      function c_get_xvec(model, mol, cn, qloc, xvec) result(stat) bind(C, name="mchrg_b200_get_xvec")
         import :: c_int, c_ptr, c_structure
         type(c_ptr), value :: model
         type(c_structure), intent(in) :: mol
         type(c_ptr), value :: cn, qloc, xvec
         integer(c_int) :: stat
      end function
      function c_get_xvec_derivs(model, mol, cn, qloc, dcndr, dcndL, dqlocdr, dqlocdL, dxdr, dxdL) &
            & result(stat) bind(C, name="mchrg_b200_get_xvec_derivs")
         import :: c_int, c_ptr, c_structure
         type(c_ptr), value :: model
         type(c_structure), intent(in) :: mol
         type(c_ptr), value :: cn, qloc, dcndr, dcndL, dqlocdr, dqlocdL, dxdr, dxdL
         integer(c_int) :: stat
      end function
      function c_get_coulomb_matrix(model, mol, cn, qloc, amat) result(stat) &
            & bind(C, name="mchrg_b200_get_coulomb_matrix")
         import :: c_int, c_ptr, c_structure
         type(c_ptr), value :: model
         type(c_structure), intent(in) :: mol
         type(c_ptr), value :: cn, qloc, amat
         integer(c_int) :: stat
      end function
      function c_get_coulomb_derivs(model, mol, cn, qloc, dcndr, dcndL, dqlocdr, dqlocdL, qvec, dadr, dadL, &
            & atrace) result(stat) bind(C, name="mchrg_b200_get_coulomb_derivs")
         import :: c_int, c_ptr, c_structure
         type(c_ptr), value :: model
         type(c_structure), intent(in) :: mol
         type(c_ptr), value :: cn, qloc, dcndr, dcndL, dqlocdr, dqlocdL, qvec, dadr, dadL, atrace
         integer(c_int) :: stat
      end function
      !> one large system over several GPUs: the library's communicator (include/mchrg_b200.h).  An MPI caller
      !> obtains the 128-byte id on rank 0 (c_comm_unique_id), MPI_Bcast's it and joins with c_comm_init_nccl;
      !> an OpenMP caller with one thread per GPU ties its contexts together with c_comm_init_local.
      function c_comm_unique_id(id) result(stat) bind(C, name="mchrg_b200_comm_unique_id")
         import :: c_ptr, c_int
         type(c_ptr), value :: id
         integer(c_int) :: stat
      end function c_comm_unique_id
      function c_comm_init_nccl(ctx, rank, nranks, id) result(stat) bind(C, name="mchrg_b200_comm_init_nccl")
         import :: c_ptr, c_int, c_int32_t
         type(c_ptr), value :: ctx
         integer(c_int32_t), value :: rank, nranks
         type(c_ptr), value :: id
         integer(c_int) :: stat
      end function c_comm_init_nccl
      function c_comm_init_local(ctxs, n) result(stat) bind(C, name="mchrg_b200_comm_init_local")
         import :: c_ptr, c_int, c_int32_t
         type(c_ptr), intent(in) :: ctxs(*)
         integer(c_int32_t), value :: n
         integer(c_int) :: stat
      end function c_comm_init_local
      function c_comm_destroy(ctx) result(stat) bind(C, name="mchrg_b200_comm_destroy")
         import :: c_ptr, c_int
         type(c_ptr), value :: ctx
         integer(c_int) :: stat
      end function c_comm_destroy
      function c_side_stream(ctx) result(stream) bind(C, name="mchrg_b200_side_stream")
         import :: c_ptr
         type(c_ptr), value :: ctx
         type(c_ptr) :: stream
      end function c_side_stream
   end interface

   !> per-solve cache payload: the B200 path rebuilds its device workspace per call, so the cache only
   !> remembers the caller's arrays for the piecewise procedures (eeq.f90:97-125)
   type :: b200_cache
      real(wp), pointer :: cn(:) => null(), qloc(:) => null()
      real(wp), pointer :: dcndr(:, :, :) => null(), dcndL(:, :, :) => null()
      real(wp), pointer :: dqlocdr(:, :, :) => null(), dqlocdL(:, :, :) => null()
   end type b200_cache

contains

!> new_eeq2019_model (src/multicharge/param.f90:46-73) on device `device`
subroutine new_b200_eeq2019_model(mol, model, error, device)
   type(structure_type), intent(in) :: mol
   class(mchrg_model_type), allocatable, intent(out) :: model
   type(error_type), allocatable, intent(out) :: error
   integer, intent(in), optional :: device
   type(b200_model_type), allocatable :: self
   integer(c_int) :: stat, dev

   dev = 0
   if (present(device)) dev = device
   allocate(self)
   stat = c_context_create(dev, self%ctx)
   if (stat == 0) stat = c_new_eeq2019_model(self%ctx, int(mol%nid, c_int32_t), int(mol%num, c_int32_t), self%handle)
   if (stat /= 0) then
      call fatal_error(error, "mchrg_b200: "//c_message())
      return
   end if
   call move_alloc(self, model)
end subroutine new_b200_eeq2019_model

!> new_eeqbc2025_model (src/multicharge/param.f90:75-125); the pairwise D3 vdW radii and Pauling electronegativities come
!> from mctc-lib on the Fortran side (get_vdw_rad, get_pauling_en: param.f90:105-115) and are passed per species (pair)
subroutine new_b200_eeqbc2025_model(mol, model, error, device)
   use mctc_data, only : get_vdw_rad, get_pauling_en
   use mctc_io_convert, only : autoaa
   type(structure_type), intent(in) :: mol
   class(mchrg_model_type), allocatable, intent(out) :: model
   type(error_type), allocatable, intent(out) :: error
   integer, intent(in), optional :: device
   type(b200_model_type), allocatable :: self
   real(c_double), allocatable :: rvdw(:, :), en(:)
   integer(c_int) :: stat, dev
   integer :: isp, jsp

   dev = 0
   if (present(device)) dev = device
   allocate(self, rvdw(mol%nid, mol%nid), en(mol%nid))
   do isp = 1, mol%nid
      en(isp) = get_pauling_en(mol%num(isp))
      do jsp = 1, mol%nid
         rvdw(jsp, isp) = get_vdw_rad(mol%num(jsp), mol%num(isp))*autoaa
      end do
   end do
   stat = c_context_create(dev, self%ctx)
   if (stat == 0) stat = c_new_eeqbc2025_model(self%ctx, int(mol%nid, c_int32_t), int(mol%num, c_int32_t), rvdw, en, &
      & self%handle)
   if (stat /= 0) then
      call fatal_error(error, "mchrg_b200: "//c_message())
      return
   end if
   call move_alloc(self, model)
end subroutine new_b200_eeqbc2025_model

!> get_charges over a batch of independent non-periodic molecules (charge.f90:37-77 per molecule): molecule m owns the atoms
!> offsets(m)+1 .. offsets(m+1); num are atomic numbers; status(m) follows the return-value convention per molecule.
!> The model must have been built for the species 1..103 (species index = atomic number).
subroutine b200_get_charges_batch(model, offsets, num, xyz, charge, qvec, status, error, energy, gradient)
   type(b200_model_type), intent(in) :: model
   integer(c_int64_t), intent(in), contiguous, target :: offsets(:)
   integer(c_int32_t), intent(in), contiguous, target :: num(:)
   real(wp), intent(in), contiguous, target :: xyz(:, :), charge(:)
   real(wp), intent(out), contiguous, target :: qvec(:)
   integer(c_int32_t), intent(out), contiguous, target :: status(:)
   type(error_type), allocatable, intent(out) :: error
   real(wp), intent(out), contiguous, optional, target :: energy(:), gradient(:, :)
   type(c_ptr) :: pe, pg
   integer(c_int) :: stat
   pe = c_null_ptr
   pg = c_null_ptr
   if (present(energy)) pe = c_loc(energy)
   if (present(gradient)) pg = c_loc(gradient)
   stat = c_singlepoint_batch(model%handle, int(size(charge), c_int32_t), c_loc(offsets), c_loc(num), c_loc(xyz), &
      & c_loc(charge), c_loc(qvec), pe, pg, c_loc(status))
   if (stat /= 0) call fatal_error(error, "mchrg_b200: "//c_message())
end subroutine b200_get_charges_batch

!> one large system on several GPUs, one MPI rank per GPU: rank 0 obtains the id, the caller broadcasts its 128 bytes
!> (MPI_Bcast) and every rank joins; from then on solve / get_charges of structures with >= 8192 atoms share the work
subroutine b200_comm_unique_id(id, error)
   character(kind=c_char), intent(out), target :: id(128)
   type(error_type), allocatable, intent(out) :: error
   if (c_comm_unique_id(c_loc(id)) /= 0) call fatal_error(error, "mchrg_b200: "//c_message())
end subroutine b200_comm_unique_id

subroutine b200_comm_init_nccl(model, rank, nranks, id, error)
   type(b200_model_type), intent(in) :: model
   integer, intent(in) :: rank, nranks
   character(kind=c_char), intent(in), target :: id(128)
   type(error_type), allocatable, intent(out) :: error
   if (c_comm_init_nccl(model%ctx, int(rank, c_int32_t), int(nranks, c_int32_t), c_loc(id)) /= 0) &
      & call fatal_error(error, "mchrg_b200: "//c_message())
end subroutine b200_comm_init_nccl

subroutine b200_free(self)
   type(b200_model_type), intent(inout) :: self
   if (c_associated(self%handle)) call c_model_free(self%handle)
   self%handle = c_null_ptr
end subroutine b200_free

function c_message() result(msg)
   character(len=:), allocatable :: msg
   character(kind=c_char), pointer :: chars(:)
   type(c_ptr) :: ptr
   integer :: i
   ptr = c_last_error()
   msg = ""
   if (.not. c_associated(ptr)) return
   call c_f_pointer(ptr, chars, [512])
   do i = 1, 512
      if (chars(i) == c_null_char) exit
      msg = msg//chars(i)
   end do
end function c_message

!> C view of a structure.  The caller's `mol` must be a TARGET (or the call chain must keep it alive and contiguous for
!> the duration of the C call: F2018 15.5.2.4 makes the addresses taken here valid only while `mol` is); `id32` is the
!> caller-owned 32-bit copy of mol%id (mol%id is a default integer, which is not c_int32_t under -fdefault-integer-8).
function to_c(mol, id32) result(cmol)
   type(structure_type), intent(in), target :: mol
   integer(c_int32_t), intent(in), target :: id32(:)
   type(c_structure) :: cmol
   cmol%nat = int(mol%nat, c_int32_t)
   cmol%nid = int(mol%nid, c_int32_t)
   cmol%id = c_loc(id32)
   cmol%xyz = c_loc(mol%xyz)
   cmol%charge = mol%charge
   cmol%periodic = merge(1_c_int32_t, 0_c_int32_t, spread(any(mol%periodic), 1, 3) .and. mol%periodic)
   cmol%lattice = 0.0_wp
   if (any(mol%periodic)) cmol%lattice = reshape(mol%lattice, [9])
end function to_c

!> c_loc of an optional contiguous array, c_null_ptr when absent
function opt3(a) result(p)
   real(wp), intent(in), contiguous, optional, target :: a(:, :, :)
   type(c_ptr) :: p
   p = c_null_ptr
   if (present(a)) p = c_loc(a)
end function opt3
function opt2(a) result(p)
   real(wp), intent(in), contiguous, optional, target :: a(:, :)
   type(c_ptr) :: p
   p = c_null_ptr
   if (present(a)) p = c_loc(a)
end function opt2
function opt1(a) result(p)
   real(wp), intent(in), contiguous, optional, target :: a(:)
   type(c_ptr) :: p
   p = c_null_ptr
   if (present(a)) p = c_loc(a)
end function opt1

!> mchrg_model_type%solve (model/type.F90:151-292): same dummy arguments, same optional protocol
subroutine b200_solve(self, mol, error, cn, qloc, dcndr, dcndL, dqlocdr, dqlocdL, &
   & energy, gradient, sigma, qvec, dqdr, dqdL)
   class(b200_model_type), intent(in) :: self
   type(structure_type), intent(in), target :: mol
   type(error_type), allocatable, intent(out) :: error
   real(wp), intent(in), contiguous, target :: cn(:)
   real(wp), intent(in), contiguous, target :: qloc(:)
   real(wp), intent(in), contiguous, optional, target :: dcndr(:, :, :)
   real(wp), intent(in), contiguous, optional, target :: dcndL(:, :, :)
   real(wp), intent(in), contiguous, optional, target :: dqlocdr(:, :, :)
   real(wp), intent(in), contiguous, optional, target :: dqlocdL(:, :, :)
   real(wp), intent(inout), contiguous, optional, target :: energy(:)
   real(wp), intent(inout), contiguous, optional, target :: gradient(:, :)
   real(wp), intent(inout), contiguous, optional, target :: sigma(:, :)
   real(wp), intent(out), contiguous, optional, target :: qvec(:)
   real(wp), intent(out), contiguous, optional, target :: dqdr(:, :, :)
   real(wp), intent(out), contiguous, optional, target :: dqdL(:, :, :)
   integer(c_int) :: stat
   integer(c_int32_t), allocatable, target :: id32(:)
   type(c_structure) :: cmol

   id32 = int(mol%id, c_int32_t)   ! 32-bit species indices that outlive the C call
   cmol = to_c(mol, id32)
   stat = c_solve(self%handle, cmol, c_loc(cn), c_loc(qloc), opt3(dcndr), opt3(dcndL), &
      & opt3(dqlocdr), opt3(dqlocdL), opt1(energy), opt2(gradient), opt2(sigma), opt1(qvec), &
      & opt3(dqdr), opt3(dqdL))
   if (stat > 0) then
      call fatal_error(error, "Bunch-Kaufman factorization failed.")  ! type.F90:223
   else if (stat < 0) then
      call fatal_error(error, "mchrg_b200: "//c_message())
   end if
end subroutine b200_solve

!> get_charges (src/multicharge/charge.f90:37-77) with the CN evaluated on the device
subroutine b200_get_charges(model, mol, error, qvec, dqdr, dqdL)
   type(b200_model_type), intent(in) :: model
   type(structure_type), intent(in), target :: mol
   type(error_type), allocatable, intent(out) :: error
   real(wp), intent(out), contiguous, target :: qvec(:)
   real(wp), intent(out), contiguous, optional, target :: dqdr(:, :, :)
   real(wp), intent(out), contiguous, optional, target :: dqdL(:, :, :)
   integer(c_int) :: stat
   integer(c_int32_t), allocatable, target :: id32(:)
   type(c_structure) :: cmol
   id32 = int(mol%id, c_int32_t)   ! 32-bit species indices that outlive the C call
   cmol = to_c(mol, id32)
   stat = c_get_charges(model%handle, cmol, c_loc(qvec), opt3(dqdr), opt3(dqdL))
   if (stat > 0) then
      call fatal_error(error, "Bunch-Kaufman factorization failed.")
   else if (stat < 0) then
      call fatal_error(error, "mchrg_b200: "//c_message())
   end if
end subroutine b200_get_charges

subroutine b200_update(self, mol, cache, cn, qloc, dcndr, dcndL, dqlocdr, dqlocdL)
   class(b200_model_type), intent(in) :: self
   type(structure_type), intent(in) :: mol
   type(cache_container), intent(inout) :: cache
   real(wp), intent(in), target :: cn(:)
   real(wp), intent(in), optional, target :: qloc(:)
   real(wp), intent(in), optional, target :: dcndr(:, :, :)
   real(wp), intent(in), optional, target :: dcndL(:, :, :)
   real(wp), intent(in), optional, target :: dqlocdr(:, :, :)
   real(wp), intent(in), optional, target :: dqlocdL(:, :, :)
   type(b200_cache), allocatable :: tmp
   allocate(tmp)
   tmp%cn => cn
   if (present(qloc)) tmp%qloc => qloc
   if (present(dcndr)) tmp%dcndr => dcndr
   if (present(dcndL)) tmp%dcndL => dcndL
   if (present(dqlocdr)) tmp%dqlocdr => dqlocdr
   if (present(dqlocdL)) tmp%dqlocdL => dqlocdL
   if (allocated(cache%raw)) deallocate(cache%raw)
   call move_alloc(tmp, cache%raw)
end subroutine b200_update

function view(cache) result(ptr)
   type(cache_container), target, intent(inout) :: cache
   type(b200_cache), pointer :: ptr
   nullify(ptr)
   select type(target => cache%raw)
   type is (b200_cache)
      ptr => target
   end select
end function view

function p1(a) result(p)
   real(wp), pointer, intent(in) :: a(:)
   type(c_ptr) :: p
   p = c_null_ptr
   if (associated(a)) p = c_loc(a)
end function p1
function p3(a) result(p)
   real(wp), pointer, intent(in) :: a(:, :, :)
   type(c_ptr) :: p
   p = c_null_ptr
   if (associated(a)) p = c_loc(a)
end function p3

subroutine b200_get_xvec(self, mol, cache, xvec)
   class(b200_model_type), intent(in) :: self
   type(structure_type), intent(in), target :: mol
   type(cache_container), intent(inout) :: cache
   real(wp), intent(out), target :: xvec(:)
   type(b200_cache), pointer :: c
   integer(c_int) :: stat
   integer(c_int32_t), allocatable, target :: id32(:)
   type(c_structure) :: cmol
   c => view(cache)
   id32 = int(mol%id, c_int32_t)   ! 32-bit species indices that outlive the C call
   cmol = to_c(mol, id32)
   stat = c_get_xvec(self%handle, cmol, p1(c%cn), p1(c%qloc), c_loc(xvec))
   if (stat /= 0) error stop "mchrg_b200: get_xvec failed"
end subroutine b200_get_xvec

subroutine b200_get_xvec_derivs(self, mol, cache, dxdr, dxdL)
   class(b200_model_type), intent(in) :: self
   type(structure_type), intent(in), target :: mol
   type(cache_container), intent(inout) :: cache
   real(wp), intent(out), contiguous, target :: dxdr(:, :, :)
   real(wp), intent(out), contiguous, target :: dxdL(:, :, :)
   type(b200_cache), pointer :: c
   integer(c_int) :: stat
   integer(c_int32_t), allocatable, target :: id32(:)
   type(c_structure) :: cmol
   c => view(cache)
   id32 = int(mol%id, c_int32_t)   ! 32-bit species indices that outlive the C call
   cmol = to_c(mol, id32)
   stat = c_get_xvec_derivs(self%handle, cmol, p1(c%cn), p1(c%qloc), p3(c%dcndr), p3(c%dcndL), &
      & p3(c%dqlocdr), p3(c%dqlocdL), c_loc(dxdr), c_loc(dxdL))
   if (stat /= 0) error stop "mchrg_b200: get_xvec_derivs failed"
end subroutine b200_get_xvec_derivs

subroutine b200_get_coulomb_matrix(self, mol, cache, amat)
   class(b200_model_type), intent(in) :: self
   type(structure_type), intent(in), target :: mol
   type(cache_container), intent(inout) :: cache
   real(wp), intent(out), target :: amat(:, :)
   type(b200_cache), pointer :: c
   integer(c_int) :: stat
   integer(c_int32_t), allocatable, target :: id32(:)
   type(c_structure) :: cmol
   c => view(cache)
   id32 = int(mol%id, c_int32_t)   ! 32-bit species indices that outlive the C call
   cmol = to_c(mol, id32)
   stat = c_get_coulomb_matrix(self%handle, cmol, p1(c%cn), p1(c%qloc), c_loc(amat))
   if (stat /= 0) error stop "mchrg_b200: get_coulomb_matrix failed"
end subroutine b200_get_coulomb_matrix

subroutine b200_get_coulomb_derivs(self, mol, cache, qvec, dadr, dadL, atrace)
   class(b200_model_type), intent(in) :: self
   type(structure_type), intent(in), target :: mol
   type(cache_container), intent(inout) :: cache
   real(wp), intent(in), target :: qvec(:)
   real(wp), intent(out), target :: dadr(:, :, :), dadL(:, :, :), atrace(:, :)
   type(b200_cache), pointer :: c
   integer(c_int) :: stat
   integer(c_int32_t), allocatable, target :: id32(:)
   type(c_structure) :: cmol
   c => view(cache)
   id32 = int(mol%id, c_int32_t)   ! 32-bit species indices that outlive the C call
   cmol = to_c(mol, id32)
   stat = c_get_coulomb_derivs(self%handle, cmol, p1(c%cn), p1(c%qloc), p3(c%dcndr), p3(c%dcndL), &
      & p3(c%dqlocdr), p3(c%dqlocdL), c_loc(qvec), c_loc(dadr), c_loc(dadL), c_loc(atrace))
   if (stat /= 0) error stop "mchrg_b200: get_coulomb_derivs failed"
end subroutine b200_get_coulomb_derivs

end module mchrg_b200_shim
