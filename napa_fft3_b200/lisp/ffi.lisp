;;; sb-alien binding of include/napafft.h.  Lisp vectors are pinned for the duration of the call
;;; (with-pinned-objects) and their storage is passed as a raw pointer (vector-sap); the library
;;; stages through its own page-locked ring, or DMAs directly if the storage was registered with
;;; napafft_host_register (e.g. a static-vectors allocation).
(in-package "NAPA-FFT.B200")

(deftype complex-sample () '(complex double-float))
(deftype complex-sample-array (&optional size) `(simple-array complex-sample (,size)))
(deftype real-sample () 'double-float)
(deftype real-sample-array (&optional size) `(simple-array real-sample (,size)))
(deftype direction () '(member 1 -1 :fwd :inv :bwd))            ; interface.lisp:3-4
(deftype scaling () '(member nil 1 t :inv sqrt :sqrt))          ; interface.lisp:6-7
(deftype windowing () '(member nil float real-sample complex complex-sample)) ; interface.lisp:9-10

(define-condition napafft-error (error)
  ((status :initarg :status :reader napafft-error-status)
   (message :initarg :message :reader napafft-error-message))
  (:report (lambda (c s)
             (format s "napafft status ~D: ~A" (napafft-error-status c) (napafft-error-message c)))))

(defun load-napafft ()
  (load-shared-object (or (sb-ext:posix-getenv "NAPAFFT_LIBRARY") "libnapafft_b200.so")))
(load-napafft)

(define-alien-routine "napafft_init" int (device int))
(define-alien-routine "napafft_last_error" c-string)
(define-alien-routine "napafft_c2c" int
  (src system-area-pointer) (dst system-area-pointer) (n size-t) (batch size-t) (stride size-t)
  (dir int) (scale int) (in-order int) (window system-area-pointer) (window-type int)
  (window-per-batch int))
(define-alien-routine "napafft_bit_reverse" int
  (src system-area-pointer) (dst system-area-pointer) (n size-t) (batch size-t) (stride size-t) (elt int))
(define-alien-routine "napafft_rfft" int
  (src system-area-pointer) (dst system-area-pointer) (n size-t) (batch size-t) (scale int))
(define-alien-routine "napafft_rifft" int
  (src system-area-pointer) (dst system-area-pointer) (n size-t) (batch size-t) (scale int))
(define-alien-routine "napafft_2rfft" int
  (v1 system-area-pointer) (v2 system-area-pointer) (dst system-area-pointer)
  (n size-t) (batch size-t) (scale int))

;; example.lisp:81-114 (filter-chunks) as one device pipeline, and complex-samplify on the device
(define-alien-routine "napafft_filter_chunks" int
  (vector system-area-pointer) (len size-t) (filter-rev system-area-pointer) (window system-area-pointer)
  (chunk size-t) (dst system-area-pointer))
(define-alien-routine "napafft_c2c_typed" int
  (src system-area-pointer) (src-type int) (dst system-area-pointer) (n size-t) (batch size-t)
  (dir int) (scale int) (in-order int) (window system-area-pointer) (window-type int)
  (window-per-batch int))

;; One transform over several GPUs (one SBCL process per GPU): include/napafft.h, "napafft_dist_*".  Rank 0 obtains the
;; 128-byte id and hands it to the other processes by whatever channel the application has (a file, a socket).
(define-alien-routine "napafft_dist_unique_id" int (out128 system-area-pointer))
(define-alien-routine "napafft_dist_init" int (unique-id128 system-area-pointer) (rank int) (nranks int))
(define-alien-routine "napafft_dist_shutdown" int)
(define-alien-routine "napafft_dist_c2c" int
  (src-local system-area-pointer) (dst-local system-area-pointer) (lgn int) (dir int))
(define-alien-routine "napafft_dist_geometry" int
  (lgn int) (nranks int) (n1 (* size-t)) (n2 (* size-t)) (cols-per-rank (* size-t)) (rows-per-rank (* size-t))
  (nchunks (* int)))

(defun check (status)
  (unless (zerop status)
    (error 'napafft-error :status status :message (napafft-last-error)))
  status)

(defun scale-code (scale)                                     ; interface.lisp:41-48
  (ecase scale ((nil 1) 0) ((t :inv) 1) ((sqrt :sqrt) 2)))
(defun direction-code (direction)                             ; interface.lisp:14-16
  (ecase direction ((1 :fwd) 1) ((-1 :inv :bwd) -1)))
(defun window-code (window)                                   ; easy-interface.lisp:39-42
  (etypecase window
    (null 0) (real-sample-array 1) (complex-sample-array 2)))

(declaim (inline element-sap))
(defun element-sap (vector index bytes-per-element)
  "Address of element INDEX of a specialised vector.  Only valid inside WITH-PINNED-OBJECTS."
  (sb-sys:sap+ (sb-sys:vector-sap vector) (* index bytes-per-element)))

(defun %c2c (src dst n start dir scale in-order window window-start)
  "Transform SRC[start, start+n) into DST[start, start+n)."
  (declare (type complex-sample-array src dst))
  (if window
      (with-pinned-objects (src dst window)
        (check (napafft-c2c (element-sap src start 16) (element-sap dst start 16)
                            n 1 n dir scale (if in-order 1 0)
                            (element-sap window window-start (if (typep window 'real-sample-array) 8 16))
                            (window-code window) 0)))
      (with-pinned-objects (src dst)
        (check (napafft-c2c (element-sap src start 16) (element-sap dst start 16)
                            n 1 n dir scale (if in-order 1 0) (sb-sys:int-sap 0) 0 0))))
  dst)
